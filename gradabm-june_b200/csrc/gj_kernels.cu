// Kernels + C ABI of the B200-native infection step (see include/gradjune_b200.h).
//
// One timestep forward = 3 passes over HBM-resident arrays:
//   F1  k_transmission      agent-major   state -> T (and quarantine-masked Tq)
//   F2  k_group_small/chunk group-major   CSR-sorted members -> per-group sums S (deterministic order)
//        (+ k_group_fix for groups that span several chunks)
//   F3  k_agent_forward     agent-major   gather S over the agent's groups -> pressure -> q ->
//                                         Gumbel-softmax draw -> state + symptoms update -> reductions
// and backward mirrors it (B1 k_agent_backward, B2 the same group kernels on cotangents, B3
// k_agent_backward_gather, k_dbeta).  No global atomics on data: group sums are segmented reductions
// over the CSR member lists in a fixed order, so results are bit-reproducible run to run.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "gj_device.cuh"

namespace gj {

static thread_local char g_err[512] = "";

static int fail(const char* what, cudaError_t e) {
  snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
  return -2;
}
static int bad(const char* what) {
  snprintf(g_err, sizeof(g_err), "invalid argument: %s", what);
  return -1;
}
#define GJ_CHECK_LAUNCH(name)                         \
  do {                                                \
    cudaError_t _e = cudaGetLastError();              \
    if (_e != cudaSuccess) return fail(name, _e);     \
  } while (0)

// ---- optional per-kernel timing with CUDA events on the launching stream (gj_profile_*) -------------
enum KernelId {
  K_TRANSMISSION = 0, K_GROUP_SMALL_F, K_GROUP_CHUNK_F, K_GROUP_FIX_F, K_AGENT_FWD,
  K_AGENT_BWD, K_GROUP_SMALL_B, K_GROUP_CHUNK_B, K_GROUP_FIX_B, K_DBETA, K_AGENT_BWD_GATHER, K_OTHER, K_COUNT
};
static const char* kKernelNames[K_COUNT] = {
  "k_transmission", "k_group_small<fwd>", "k_group_chunk<fwd>", "k_group_fix<fwd>", "k_agent_forward",
  "k_agent_backward", "k_group_small<bwd>", "k_group_chunk<bwd>", "k_group_fix<bwd>", "k_dbeta",
  "k_agent_backward_gather", "other"};
constexpr int kMaxProfiled = 16384;
struct Profiler {
  bool on = false;
  int n = 0;
  int64_t launches[K_COUNT] = {0};
  cudaEvent_t ev[kMaxProfiled][2];
  int id[kMaxProfiled];
  bool created = false;
};
static Profiler g_prof;
struct ProfScope {
  int slot = -1;
  cudaStream_t st;
  ProfScope(int kid, cudaStream_t s) : st(s) {
    g_prof.launches[kid]++;
    if (g_prof.on && g_prof.n < kMaxProfiled) {
      slot = g_prof.n++;
      g_prof.id[slot] = kid;
      cudaEventRecord(g_prof.ev[slot][0], st);
    }
  }
  ~ProfScope() {
    if (slot >= 0) cudaEventRecord(g_prof.ev[slot][1], st);
  }
};

static inline int blocks_for(int64_t n, int per_block) {
  int64_t b = (n + per_block - 1) / per_block;
  return (int)(b < 1 ? 1 : b);
}
static inline int agent_grid(int64_t n) {
  int64_t b = (n + kBlock - 1) / kBlock;
  if (b > kRedBlocks) b = kRedBlocks;
  return (int)(b < 1 ? 1 : b);
}

// scratch layout (bytes): [0,128) tickets | red partials double[kRedBlocks][kMaxRed] |
//                         dbeta partials double[GJ_MAX_NETS][kRedBlocks] | part_a float[n_parts*nets] | part_b ...
struct Scratch {
  unsigned int* tickets;
  double* red_part;
  double* dbeta_part;
  float* part_a;  // [GJ_MAX_NETS][n_parts]
  float* part_b;
};
static inline int64_t scratch_bytes(const gj_world_desc* w) {
  int64_t b = 128;
  b += (int64_t)sizeof(double) * kRedBlocks * kMaxRed;
  b += (int64_t)sizeof(double) * GJ_MAX_NETS * kRedBlocks;
  b += 2 * (int64_t)sizeof(float) * GJ_MAX_NETS * (w->n_parts > 0 ? w->n_parts : 1);
  return (b + 255) / 256 * 256;
}
static inline Scratch carve(const gj_world_desc* w, void* base) {
  Scratch s;
  char* p = (char*)base;
  s.tickets = (unsigned int*)p;
  p += 128;
  s.red_part = (double*)p;
  p += sizeof(double) * kRedBlocks * kMaxRed;
  s.dbeta_part = (double*)p;
  p += sizeof(double) * GJ_MAX_NETS * kRedBlocks;
  s.part_a = (float*)p;
  p += sizeof(float) * GJ_MAX_NETS * (w->n_parts > 0 ? w->n_parts : 1);
  s.part_b = (float*)p;
  return s;
}

// per-type channel table derived from the net list (host, passed by value)
struct Channels {
  int nch[GJ_MAX_TYPES];
  int net[GJ_MAX_TYPES][GJ_MAX_CHANNELS];
};

// ================================================================================================
// F1  transmission
// ================================================================================================
__global__ void __launch_bounds__(kBlock) k_transmission(int64_t n, float now, const float* __restrict__ tinf,
                                                         const float* __restrict__ inf,
                                                         const float* __restrict__ maxinf,
                                                         const float* __restrict__ shape,
                                                         const float* __restrict__ rate,
                                                         const float* __restrict__ shift,
                                                         const float* __restrict__ k0, float* __restrict__ T,
                                                         // optional quarantine-masked copy
                                                         gj_step_params p, const float* __restrict__ cur,
                                                         float* __restrict__ Tq) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; a < n; a += stride) {
    const TransTerms tt = transmission_terms<false>(now, tinf[a], maxinf[a], shape[a], rate[a], shift[a], k0[a]);
    const float t = tt.coef * inf[a];
    T[a] = t;
    if (Tq != nullptr && Tq != T) Tq[a] = quarantine_mask(p, cur[a]) * t;
  }
}

__global__ void __launch_bounds__(kBlock) k_mask_transmission(int64_t n, gj_step_params p, const float* __restrict__ cur,
                                                              const float* __restrict__ T, float* __restrict__ Tq) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; a < n; a += stride)
    Tq[a] = quarantine_mask(p, cur[a]) * T[a];
}

__global__ void __launch_bounds__(kBlock) k_transmission_bwd(int64_t n, float now, const float* __restrict__ tinf,
                                                             const float* __restrict__ inf,
                                                             const float* __restrict__ maxinf,
                                                             const float* __restrict__ shape,
                                                             const float* __restrict__ rate,
                                                             const float* __restrict__ shift,
                                                             const float* __restrict__ k0,
                                                             const float* __restrict__ gT, float* __restrict__ g_tinf,
                                                             float* __restrict__ g_inf) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; a < n; a += stride) {
    const TransTerms tt = transmission_terms<true>(now, tinf[a], maxinf[a], shape[a], rate[a], shift[a], k0[a]);
    const float g = gT[a];
    if (g_inf) g_inf[a] = g * tt.coef;
    if (g_tinf) g_tinf[a] = g * (tt.dcoef * inf[a]);
  }
}

__global__ void __launch_bounds__(kBlock) k_profile_prepare(int64_t n, const float* __restrict__ shape,
                                                            float* __restrict__ k0) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; a < n; a += stride)
    k0[a] = expf(-lgammaf(shape[a]));
}

// ================================================================================================
// F2 / B2  group-major segmented sums
//   forward : in0 = T,  in1 = Tq -> out_a = S_scaled (sum of v*c_g), out_b = S_unscaled (sum of v)
//   backward: in0 = w,  in1 = wq -> out_a = cR (c_g * sum v),        out_b = R (sum v)
// v per member and channel (network sharing the edge type):
//   HOUSEHOLD in0 ; PLAIN in1 ; LEISURE p*in1 ; CARE_VISIT p*in1 (forward) / p*in1*(age>75) (backward)
// ================================================================================================
struct GroupAcc {
  float a[GJ_MAX_CHANNELS];
  float b[GJ_MAX_CHANNELS];
};

template <bool kBwd>
__device__ __forceinline__ void group_member(const gj_step_params& p, const Channels& ch, int type, int nch,
                                             const float* __restrict__ beta_c, float pcg, uint32_t agent,
                                             const float* __restrict__ in0, const float* __restrict__ in1,
                                             const uint8_t* __restrict__ cls, const float* __restrict__ lprob,
                                             GroupAcc& acc) {
  const float v0 = in0[agent];
  const float v1 = (in1 == in0) ? v0 : in1[agent];
  int c8 = -1;
#pragma unroll
  for (int c = 0; c < GJ_MAX_CHANNELS; ++c) {
    if (c < nch) {
      const gj_net& net = p.nets[ch.net[type][c]];
      float v;
      if (net.kind == GJ_KIND_HOUSEHOLD) {
        v = v0;
      } else if (net.kind == GJ_KIND_PLAIN) {
        v = v1;
      } else {
        if (c8 < 0) c8 = cls[agent];
        v = leisure_prob(lprob, net.prob_row, p.day_type, c8) * v1;
        if (kBwd && net.kind == GJ_KIND_CARE_VISIT) v = v * (((c8 % 100) > 75) ? 1.0f : 0.0f);
      }
      if (!kBwd) acc.a[c] += v * (beta_c[c] * pcg);  // message = T' * (beta * p_contact)   base.py:70,86-87
      acc.b[c] += v;
    }
  }
}

template <bool kBwd>
__device__ __forceinline__ void group_store(const gj_world_desc& w, const gj_step_params& p, const Channels& ch,
                                            int type, int nch, const float* __restrict__ beta_c, float pcg, uint32_t g,
                                            const GroupAcc& acc, float* __restrict__ out_a,
                                            float* __restrict__ out_b) {
  const int64_t lg = (int64_t)g - w.type_group_off[type];
#pragma unroll
  for (int c = 0; c < GJ_MAX_CHANNELS; ++c) {
    if (c < nch) {
      const int64_t o = (int64_t)p.nets[ch.net[type][c]].s_off + lg;
      out_a[o] = kBwd ? (beta_c[c] * pcg) * acc.b[c] : acc.a[c];
      out_b[o] = acc.b[c];
    }
  }
}

template <bool kBwd>
__global__ void __launch_bounds__(kBlock) k_group_small(gj_world_desc w, gj_step_params p, Channels ch,
                                                        const float* __restrict__ beta,
                                                        const float* __restrict__ lprob,
                                                        const float* __restrict__ in0,
                                                        const float* __restrict__ in1, float* __restrict__ out_a,
                                                        float* __restrict__ out_b) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= w.n_small) return;
  const uint32_t g = w.small_groups[i];
  const int type = type_of_group(w, g);
  const int nch = ch.nch[type];
  if (nch == 0) return;
  float beta_c[GJ_MAX_CHANNELS];
#pragma unroll
  for (int c = 0; c < GJ_MAX_CHANNELS; ++c) beta_c[c] = (c < nch) ? beta[ch.net[type][c]] : 0.0f;
  const float pcg = w.pc[g];
  GroupAcc acc;
#pragma unroll
  for (int c = 0; c < GJ_MAX_CHANNELS; ++c) acc.a[c] = acc.b[c] = 0.0f;
  const uint32_t b = w.gm_ptr[g], e = w.gm_ptr[g + 1];
  for (uint32_t j = b; j < e; ++j)  // sequential in the reference's edge order (bit-exact small groups)
    group_member<kBwd>(p, ch, type, nch, beta_c, pcg, w.gm_agent[j], in0, in1, w.cls, lprob, acc);
  group_store<kBwd>(w, p, ch, type, nch, beta_c, pcg, g, acc, out_a, out_b);
}

// one warp per chunk (<= GJ_CHUNK members): lanes stride over the member list, then a fixed shuffle tree
template <bool kBwd>
__global__ void __launch_bounds__(kBlock) k_group_chunk(gj_world_desc w, gj_step_params p, Channels ch,
                                                        const float* __restrict__ beta,
                                                        const float* __restrict__ lprob,
                                                        const float* __restrict__ in0,
                                                        const float* __restrict__ in1, float* __restrict__ out_a,
                                                        float* __restrict__ out_b, float* __restrict__ part_a,
                                                        float* __restrict__ part_b) {
  const int lane = threadIdx.x & 31;
  const int64_t ci = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (ci >= w.n_chunks) return;
  const uint32_t g = w.chunk_group[ci];
  const int type = type_of_group(w, g);
  const int nch = ch.nch[type];
  if (nch == 0) return;
  float beta_c[GJ_MAX_CHANNELS];
#pragma unroll
  for (int c = 0; c < GJ_MAX_CHANNELS; ++c) beta_c[c] = (c < nch) ? beta[ch.net[type][c]] : 0.0f;
  const float pcg = w.pc[g];
  GroupAcc acc;
#pragma unroll
  for (int c = 0; c < GJ_MAX_CHANNELS; ++c) acc.a[c] = acc.b[c] = 0.0f;
  const uint32_t b = w.chunk_begin[ci], e = w.chunk_end[ci];
  for (uint32_t j = b + lane; j < e; j += 32)
    group_member<kBwd>(p, ch, type, nch, beta_c, pcg, w.gm_agent[j], in0, in1, w.cls, lprob, acc);
#pragma unroll
  for (int c = 0; c < GJ_MAX_CHANNELS; ++c) {
    if (c < nch) {
      if (!kBwd) acc.a[c] = warp_sum(acc.a[c]);
      acc.b[c] = warp_sum(acc.b[c]);
    }
  }
  if (lane != 0) return;
  const int part = w.chunk_part[ci];
  if (part < 0) {
    group_store<kBwd>(w, p, ch, type, nch, beta_c, pcg, g, acc, out_a, out_b);
  } else {
#pragma unroll
    for (int c = 0; c < GJ_MAX_CHANNELS; ++c) {
      if (c < nch) {
        part_a[(int64_t)c * w.n_parts + part] = acc.a[c];
        part_b[(int64_t)c * w.n_parts + part] = acc.b[c];
      }
    }
  }
}

// groups spanning several chunks: add the chunk partials in chunk order
template <bool kBwd>
__global__ void __launch_bounds__(kBlock) k_group_fix(gj_world_desc w, gj_step_params p, Channels ch,
                                                      const float* __restrict__ beta,
                                                      const float* __restrict__ part_a,
                                                      const float* __restrict__ part_b, float* __restrict__ out_a,
                                                      float* __restrict__ out_b) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= w.n_big) return;
  const uint32_t g = w.big_groups[i];
  const int type = type_of_group(w, g);
  const int nch = ch.nch[type];
  if (nch == 0) return;
  float beta_c[GJ_MAX_CHANNELS];
#pragma unroll
  for (int c = 0; c < GJ_MAX_CHANNELS; ++c) beta_c[c] = (c < nch) ? beta[ch.net[type][c]] : 0.0f;
  GroupAcc acc;
#pragma unroll
  for (int c = 0; c < GJ_MAX_CHANNELS; ++c) acc.a[c] = acc.b[c] = 0.0f;
  for (uint32_t j = w.big_part_ptr[i]; j < w.big_part_ptr[i + 1]; ++j) {
#pragma unroll
    for (int c = 0; c < GJ_MAX_CHANNELS; ++c) {
      if (c < nch) {
        acc.a[c] += part_a[(int64_t)c * w.n_parts + j];
        acc.b[c] += part_b[(int64_t)c * w.n_parts + j];
      }
    }
  }
  group_store<kBwd>(w, p, ch, type, nch, beta_c, w.pc[g], g, acc, out_a, out_b);
}

// ================================================================================================
// block reduction of a few doubles + "last block finishes" (fixed summation order -> deterministic)
// ================================================================================================
template <int kR>
__device__ __forceinline__ void block_reduce_finish(double (&v)[kR], int nr, double* __restrict__ partials,
                                                    unsigned int* __restrict__ ticket, float* __restrict__ out) {
  __shared__ double sm[kBlock / 32][kR];
  __shared__ bool last;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int r = 0; r < kR; ++r)
    if (r < nr) v[r] = warp_sum(v[r]);
  if (lane == 0)
    for (int r = 0; r < nr; ++r) sm[wid][r] = v[r];
  __syncthreads();
  if (threadIdx.x < nr) {
    double s = 0.0;
    for (int k = 0; k < kBlock / 32; ++k) s += sm[k][threadIdx.x];
    partials[(int64_t)blockIdx.x * kR + threadIdx.x] = s;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(ticket, 1u);
    last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (last) {
    __threadfence();
    if (threadIdx.x < nr) {
      double s = 0.0;
      for (unsigned int b = 0; b < gridDim.x; ++b) s += partials[(int64_t)b * kR + threadIdx.x];
      out[threadIdx.x] = (float)s;
    }
    if (threadIdx.x == 0) *ticket = 0u;  // leave the scratch zeroed for the next launch
  }
}

// ================================================================================================
// F3  agent-major forward
// ================================================================================================
struct Masks {
  float mT, mS, mS_age;  // mS_age: care-visit (age > 75) factor
};

__device__ __forceinline__ Masks net_masks(const gj_step_params& p, const gj_net& net, float mq, int cls,
                                           const float* __restrict__ lprob) {
  Masks m;
  m.mS_age = 1.0f;
  if (net.kind == GJ_KIND_HOUSEHOLD) {
    m.mT = m.mS = 1.0f;
  } else if (net.kind == GJ_KIND_PLAIN) {
    m.mT = m.mS = mq;
  } else {
    const float lm = leisure_prob(lprob, net.prob_row, p.day_type, cls);
    m.mT = m.mS = mq * lm;
    if (net.kind == GJ_KIND_CARE_VISIT) m.mS_age = ((cls % 100) > 75) ? 1.0f : 0.0f;
  }
  return m;
}

__global__ void __launch_bounds__(kBlock) k_agent_forward(gj_world_desc w, gj_step_params p, gj_fwd_io io,
                                                          double* __restrict__ red_part,
                                                          unsigned int* __restrict__ ticket) {
  const int64_t N = w.n_agents;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int dead = p.n_stages - 1;
  double red[kMaxRed];
#pragma unroll
  for (int r = 0; r < kMaxRed; ++r) red[r] = 0.0;

  for (int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; a < N; a += stride) {
    const int cls = w.cls ? w.cls[a] : 0;
    const int age = cls % 100;
    float s = io.s ? io.s[a] : 0.0f;
    float inf = io.inf ? io.inf[a] : 0.0f;
    float tinf = io.tinf ? io.tinf[a] : 0.0f;
    float cur = io.cur ? io.cur[a] : 1.0f;
    float nxt = io.nxt ? io.nxt[a] : 1.0f;
    float ttn = io.ttn ? io.ttn[a] : 0.0f;

    float q = 1.0f;
    // ---- InfectionNetworks.forward: gather the group sums of every active network ----------------
    if (p.phases & GJ_PHASE_NETWORKS) {
      const float mq = (p.n_quar > 0) ? quarantine_mask(p, cur) : 1.0f;
      float lam = 0.0f, X = 0.0f;
      const uint32_t e0 = w.am_ptr[a], e1 = w.am_ptr[a + 1];
      for (int k = 0; k < p.n_nets; ++k) {
        const gj_net net = p.nets[k];
        const Masks m = net_masks(p, net, mq, cls, io.leisure_prob);
        float sp = m.mS * s;  // susceptibilities = mask * [leisure_mask *] susceptibility
        float sx = m.mS;
        if (net.kind == GJ_KIND_CARE_VISIT) {
          sp = sp * m.mS_age;
          sx = sx * m.mS_age;
        }
        float Pk = 0.0f, PXk = 0.0f;
        for (uint32_t j = e0; j < e1; ++j) {
          const uint32_t ent = w.am_ent[j];
          if ((int)(ent >> 28) == net.type) {
            const float Sg = io.S_scaled[(int64_t)net.s_off + (ent & 0x0FFFFFFFu)];
            Pk += Sg * sp;  // message = cumulative_trans * susceptibility   base.py:80-87
            PXk += Sg * sx;
          }
        }
        lam += Pk;  // trans_susc += network(...)   base.py:133-135
        X += PXk;
      }
      q = not_infected_prob(lam, p.dt);
      if (io.tape_v) io.tape_v[a] = (s == 0.0f) ? X : lam;
      if (io.q) io.q[a] = q;
      if (io.lam) io.lam[a] = lam;
    } else if (io.q_in) {
      q = io.q_in[a];
    }
    if (p.mode == GJ_MODE_SEED) {
      const float f = io.seed_fraction[0];
      const float probs = f * 1.0f;
      q = 1.0f - probs;  // infection.py:36-40
    }

    // ---- IsInfectedSampler.forward -----------------------------------------------------------------
    float n = 0.0f;
    StepNoise nz;
    nz.E0 = nz.E1 = 1.0f;
    nz.u = 0.0f;
    const bool need_noise = (p.phases & (GJ_PHASE_SAMPLE | GJ_PHASE_SYMPTOMS)) != 0;
    if (need_noise) {
      if (io.inj_E == nullptr || io.inj_u == nullptr) nz = draw_step_noise(p.seed, p.call_index, a);
      if (io.inj_E) {
        nz.E0 = io.inj_E[a];
        nz.E1 = io.inj_E[N + a];
      }
      if (io.inj_u) nz.u = io.inj_u[a];
    }
    if (p.phases & GJ_PHASE_SAMPLE) {
      const Draw d = gumbel_draw(q, nz.E0, nz.E1, p.tau);
      n = d.n;
      if (io.tape_y0) io.tape_y0[a] = d.ty;
    } else if (io.n_in) {
      n = io.n_in[a];
    }
    if (io.n) io.n[a] = n;

    // ---- infect_people ---------------------------------------------------------------------------
    if (p.phases & GJ_PHASE_INFECT) {
      s = fmaxf(0.0f, s - n);  // maximum(0, s - n) and clamp(s - n, min=0) agree in value
      inf = inf + n;
      tinf = tinf + n * (p.now - tinf);
      if (io.s_o) io.s_o[a] = s;
      if (io.inf_o) io.inf_o[a] = inf;
      if (io.tinf_o) io.tinf_o[a] = tinf;
    }

    // ---- SymptomsUpdater.forward -----------------------------------------------------------------
    if (p.phases & GJ_PHASE_SYMPTOMS) {
      const float* inj_z = io.inj_z;
      const uint64_t seed = p.seed;
      const uint32_t call = p.call_index;
      const SympOut so = symptoms_forward(p, io.stage_prob, cur, nxt, ttn, n, age, nz.u, [&](int row) {
        return inj_z ? inj_z[(int64_t)row * N + a] : draw_step_normal(seed, call, a);
      });
      cur = so.cur;
      nxt = so.nxt;
      ttn = so.ttn;
      if (io.cur_o) io.cur_o[a] = cur;
      if (io.nxt_o) io.nxt_o[a] = nxt;
      if (io.ttn_o) io.ttn_o[a] = ttn;
    }

    // ---- Runner.forward reductions (runner.py:167-171,198-224) ------------------------------------
    if (io.red) {
      red[0] += (double)inf;
      red[1] += (cur == (float)dead) ? (double)(cur / (float)dead) : 0.0;
      for (int b = 0; b < p.n_age_bins; ++b)
        if (age > p.age_bins[b] && age < p.age_bins[b + 1]) red[2 + b] += (double)inf;
    }
  }
  if (io.red) block_reduce_finish<kMaxRed>(red, 2 + p.n_age_bins, red_part, ticket, io.red);
}

// ================================================================================================
// B1  agent-major backward, part 1: symptoms^T, infect^T, sampler^T, pressure^T up to the group sums
// ================================================================================================
__global__ void __launch_bounds__(kBlock) k_agent_backward(gj_world_desc w, gj_step_params p, gj_bwd_io io,
                                                           double* __restrict__ red_part,
                                                           unsigned int* __restrict__ ticket) {
  const int64_t N = w.n_agents;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int dead = p.n_stages - 1;
  double gfrac[1] = {0.0};
  const bool seed_mode = p.mode == GJ_MODE_SEED;

  for (int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; a < N; a += stride) {
    const int cls = w.cls ? w.cls[a] : 0;
    const int age = cls % 100;
    const float s = io.s ? io.s[a] : 0.0f;
    const float inf = io.inf ? io.inf[a] : 0.0f;
    const float tinf = io.tinf ? io.tinf[a] : 0.0f;
    const float cur = io.cur ? io.cur[a] : 1.0f;
    const float nxt = io.nxt ? io.nxt[a] : 1.0f;
    const float ttn = io.ttn ? io.ttn[a] : 0.0f;
    float n = 0.0f;
    if (io.inf_o) n = io.inf_o[a] - inf;  // exact: both are small integers
    else if (io.n_in) n = io.n_in[a];

    float gs_o = io.g_s_o ? io.g_s_o[a] : 0.0f;
    float ginf_o = io.g_inf_o ? io.g_inf_o[a] : 0.0f;
    float gtinf_o = io.g_tinf_o ? io.g_tinf_o[a] : 0.0f;
    float gcur_o = io.g_cur_o ? io.g_cur_o[a] : 0.0f;
    float gnxt_o = io.g_nxt_o ? io.g_nxt_o[a] : 0.0f;
    float gttn_o = io.g_ttn_o ? io.g_ttn_o[a] : 0.0f;
    float gn = io.g_n ? io.g_n[a] : 0.0f;

    float gcur = gcur_o, gnxt = gnxt_o, gttn = gttn_o;
    // ---- symptoms^T ------------------------------------------------------------------------------
    if (p.phases & GJ_PHASE_SYMPTOMS) {
      float u = 0.0f;
      if (io.inj_u) u = io.inj_u[a];
      else u = draw_step_noise(p.seed, p.call_index, a).u;
      const float* inj_z = io.inj_z;
      const uint64_t seed = p.seed;
      const uint32_t call = p.call_index;
      const SympOut so = symptoms_forward(p, io.stage_prob, cur, nxt, ttn, n, age, u, [&](int row) {
        return inj_z ? inj_z[(int64_t)row * N + a] : draw_step_normal(seed, call, a);
      });
      // reductions fold in here: deaths = sum (cur' == dead) * cur' / dead   runner.py:204-209
      if (io.g_red && so.cur == (float)dead) gcur_o += io.g_red[1] / (float)dead;
      float gcur1 = gcur_o;
      float gnxt1 = gnxt_o;
      const float gttn1 = gttn_o;
      if (so.branch == 1) {         // next += m ; ttn += dwell * m ; m = (cur==i)*cur/i * transition * symp
        gcur1 += (gnxt_o + gttn_o * so.dwell) / (float)so.stage;
      } else if (so.branch == 2) {  // next -= next * m ; ttn += dwell * m
        gcur1 += (gttn_o * so.dwell - gnxt_o * so.nxt1) / (float)so.stage;
        gnxt1 = 0.0f;               // d(next - next*m)/dnext = 1 - m = 0
      }
      gcur = gcur1 * (1.0f - so.tr);  // cur' = cur - (cur - next1) * transition
      gnxt1 += gcur1 * so.tr;
      gnxt = gnxt1 * (1.0f - n);      // next1 = next + n * (2 - next)
      gn += gnxt1 * (2.0f - nxt);
      gttn = gttn1 * (1.0f - n);      // ttn1 = ttn + n * (now - ttn)
      gn += gttn1 * (p.now - ttn);
    } else if (io.g_red && cur == (float)dead) {
      gcur += io.g_red[1] / (float)dead;
    }

    // ---- infect^T -------------------------------------------------------------------------------
    float gs = gs_o, ginf = ginf_o, gtinf = gtinf_o;
    if (io.g_red) {
      ginf_o += io.g_red[0];
      for (int b = 0; b < p.n_age_bins; ++b)
        if (age > p.age_bins[b] && age < p.age_bins[b + 1]) ginf_o += io.g_red[2 + b];
      ginf = ginf_o;
    }
    if (p.phases & GJ_PHASE_INFECT) {
      const float d = s - n;
      float wgt;
      if (seed_mode) wgt = (d >= 0.0f) ? 1.0f : 0.0f;            // clamp(min=0): gradient where x >= min
      else wgt = (d > 0.0f) ? 1.0f : ((d == 0.0f) ? 0.5f : 0.0f);  // maximum(0, x): ties split 1/2
      gs = gs_o * wgt;
      gn += -(gs_o * wgt) + ginf_o + gtinf_o * (p.now - tinf);
      ginf = ginf_o;
      gtinf = gtinf_o * (1.0f - n);
    }

    // ---- sampler^T -------------------------------------------------------------------------------
    float gq = io.g_q ? io.g_q[a] : 0.0f;
    float lam = 0.0f, q = 1.0f, v = 0.0f;
    if (p.phases & GJ_PHASE_NETWORKS) {
      v = io.tape_v[a];
      lam = (s == 0.0f) ? 0.0f : v;
      q = not_infected_prob(lam, p.dt);
    }
    if (seed_mode) q = 1.0f - io.seed_fraction[0] * 1.0f;
    if (p.phases & GJ_PHASE_SAMPLE) {
      if (!(p.phases & GJ_PHASE_NETWORKS) && !seed_mode && io.q_in) q = io.q_in[a];  // stand-alone sampler
      float y0, y1;
      decode_soft(io.tape_y0[a], y0, y1);
      const float gret0 = -gn;                        // new_infected = 1 - ret[0]
      const float dot = gret0 * y0;                   // softmax^T: (g - sum(g*y)) * y with g = (gret0, 0)
      const float gx0 = (gret0 - dot) * y0;
      const float gx1 = (0.0f - dot) * y1;
      const float gl0 = gx0 / p.tau, gl1 = gx1 / p.tau;
      gq += gl0 / q - gl1 / (1.0f - q);               // logits = log([q, 1-q])
    }
    if (io.g_q_out) io.g_q_out[a] = gq;
    if (io.g_n_out) io.g_n_out[a] = gn;
    if (seed_mode) gfrac[0] += (double)(-gq);     // q = 1 - fraction

    // ---- pressure^T, agent side ------------------------------------------------------------------
    if (p.phases & GJ_PHASE_NETWORKS) {
      // q = clamp(exp(-clamp(lam)*dt), 0, 1): clamp passes the gradient on its closed interval
      float glam = 0.0f;
      if (q >= 0.0f && q <= 1.0f) {
        const float glc = gq * q * (-p.dt);
        if (lam >= 1e-6f && lam <= 100.0f) glam = glc;
      }
      if (io.g_lam) glam += io.g_lam[a];
      const float X = (s == 0.0f) ? v : v / s;
      gs += glam * X;
      const float mq = (p.n_quar > 0) ? quarantine_mask(p, cur) : 1.0f;
      const float wv = glam * s;
      io.w[a] = wv;
      if (io.wq != io.w) io.wq[a] = glam * (mq * s);
    }
    if (io.g_s) io.g_s[a] = gs;
    if (io.g_inf) io.g_inf[a] = ginf;
    if (io.g_tinf) io.g_tinf[a] = gtinf;
    if (io.g_cur) io.g_cur[a] = gcur;
    if (io.g_nxt) io.g_nxt[a] = gnxt;
    if (io.g_ttn) io.g_ttn[a] = gttn;
  }
  if (seed_mode && io.g_seed_fraction) block_reduce_finish<1>(gfrac, 1, red_part, ticket, io.g_seed_fraction);
}

// ================================================================================================
// B3  agent-major backward, part 2: gather c_g * R_g -> dL/dT -> (is_infected, infection_time)
// ================================================================================================
__global__ void __launch_bounds__(kBlock) k_agent_backward_gather(gj_world_desc w, gj_step_params p, gj_bwd_io io) {
  const int64_t N = w.n_agents;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; a < N; a += stride) {
    const int cls = w.cls[a];
    const float cur = io.cur ? io.cur[a] : 1.0f;
    const float mq = (p.n_quar > 0) ? quarantine_mask(p, cur) : 1.0f;
    float gT = 0.0f;
    const uint32_t e0 = w.am_ptr[a], e1 = w.am_ptr[a + 1];
    for (int k = 0; k < p.n_nets; ++k) {
      const gj_net net = p.nets[k];
      const Masks m = net_masks(p, net, mq, cls, io.leisure_prob);
      float acc = 0.0f;
      for (uint32_t j = e0; j < e1; ++j) {
        const uint32_t ent = w.am_ent[j];
        if ((int)(ent >> 28) == net.type) acc += io.cR[(int64_t)net.s_off + (ent & 0x0FFFFFFFu)];
      }
      gT += m.mT * acc;
    }
    if (io.g_T) {
      io.g_T[a] = gT;
    } else {
      const TransTerms tt =
          transmission_terms<true>(p.now, io.tinf[a], io.maxinf[a], io.shape[a], io.rate[a], io.shift[a], io.k0[a]);
      if (io.g_inf) io.g_inf[a] += gT * tt.coef;
      if (io.g_tinf) io.g_tinf[a] += gT * (tt.dcoef * io.inf[a]);
    }
  }
}

// dL/dbeta_k = sum_g pc_g * S~_g * R_g   (fixed-order two-level sum in fp64)
__global__ void __launch_bounds__(kBlock) k_dbeta(gj_world_desc w, gj_step_params p, const float* __restrict__ S_un,
                                                  const float* __restrict__ R, double* __restrict__ partials,
                                                  unsigned int* __restrict__ tickets, float* __restrict__ g_beta) {
  const int k = blockIdx.y;
  const gj_net net = p.nets[k];
  const int64_t g0 = w.type_group_off[net.type];
  const int64_t G = w.type_group_off[net.type + 1] - g0;
  double acc[1] = {0.0};
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < G; i += (int64_t)gridDim.x * blockDim.x)
    acc[0] += (double)(w.pc[g0 + i] * S_un[(int64_t)net.s_off + i]) * (double)R[(int64_t)net.s_off + i];
  block_reduce_finish<1>(acc, 1, partials + (int64_t)k * kRedBlocks, tickets + 1 + k, g_beta + k);
}

__global__ void k_philox_fill(uint64_t seed, uint32_t call, int64_t n, float* __restrict__ E, float* __restrict__ u,
                              float* __restrict__ z) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; a < n; a += stride) {
    const StepNoise nz = draw_step_noise(seed, call, a);
    if (E) {
      E[a] = nz.E0;
      E[n + a] = nz.E1;
    }
    if (u) u[a] = nz.u;
    if (z) z[a] = draw_step_normal(seed, call, a);
  }
}

// ================================================================================================
// host side
// ================================================================================================
static int build_channels(const gj_world_desc* w, const gj_step_params* p, Channels* ch) {
  memset(ch, 0, sizeof(*ch));
  if (p->n_nets < 0 || p->n_nets > GJ_MAX_NETS) return bad("n_nets");
  for (int k = 0; k < p->n_nets; ++k) {
    const int t = p->nets[k].type;
    if (t < 0 || t >= w->n_types) return bad("net.type");
    if (ch->nch[t] >= GJ_MAX_CHANNELS) return bad("too many networks share one edge type");
    ch->net[t][ch->nch[t]++] = k;
  }
  return 0;
}

static int check_world(const gj_world_desc* w) {
  if (!w) return bad("world is NULL");
  if (w->n_agents < 0 || w->n_types < 0 || w->n_types > GJ_MAX_TYPES) return bad("world sizes");
  if (w->n_edges >= ((int64_t)1 << 32) || w->n_agents >= ((int64_t)1 << 32)) return bad("world too large for 32-bit CSR");
  for (int t = 0; t < w->n_types; ++t)
    if (w->type_group_off[t + 1] - w->type_group_off[t] >= ((int64_t)1 << 28)) return bad("too many groups in one type");
  return 0;
}

template <bool kBwd>
static int launch_group_pass(const gj_world_desc* w, const gj_step_params* p, const Channels& ch, const float* beta,
                             const float* lprob, const float* in0, const float* in1, float* out_a, float* out_b,
                             const Scratch& sc, cudaStream_t st) {
  if (w->n_small > 0) {
    ProfScope ps(kBwd ? K_GROUP_SMALL_B : K_GROUP_SMALL_F, st);
    k_group_small<kBwd><<<blocks_for(w->n_small, kBlock), kBlock, 0, st>>>(*w, *p, ch, beta, lprob, in0, in1, out_a,
                                                                         out_b);
    GJ_CHECK_LAUNCH("k_group_small");
  }
  if (w->n_chunks > 0) {
    ProfScope ps(kBwd ? K_GROUP_CHUNK_B : K_GROUP_CHUNK_F, st);
    k_group_chunk<kBwd><<<blocks_for(w->n_chunks * 32, kBlock), kBlock, 0, st>>>(*w, *p, ch, beta, lprob, in0, in1,
                                                                              out_a, out_b, sc.part_a, sc.part_b);
    GJ_CHECK_LAUNCH("k_group_chunk");
  }
  if (w->n_big > 0) {
    ProfScope ps(kBwd ? K_GROUP_FIX_B : K_GROUP_FIX_F, st);
    k_group_fix<kBwd><<<blocks_for(w->n_big, kBlock), kBlock, 0, st>>>(*w, *p, ch, beta, sc.part_a, sc.part_b, out_a,
                                                                     out_b);
    GJ_CHECK_LAUNCH("k_group_fix");
  }
  return 0;
}

}  // namespace gj

using namespace gj;

extern "C" {

int gj_abi_version(void) { return GJ_ABI_VERSION; }
const char* gj_last_error(void) { return g_err; }

int gj_config(int64_t* out, int n) {
  const int64_t v[7] = {GJ_SMALL_GROUP,      GJ_CHUNK,          (int64_t)sizeof(gj_world_desc), (int64_t)sizeof(gj_step_params),
                        (int64_t)sizeof(gj_fwd_io), (int64_t)sizeof(gj_bwd_io), kRedBlocks};
  for (int i = 0; i < n && i < 7; ++i) out[i] = v[i];
  return 7;
}

int64_t gj_scratch_bytes(const gj_world_desc* w) { return w ? scratch_bytes(w) : -1; }

int gj_profile_enable(int on) {
  if (on && !g_prof.created) {
    for (int i = 0; i < kMaxProfiled; ++i) {
      if (cudaEventCreate(&g_prof.ev[i][0]) != cudaSuccess || cudaEventCreate(&g_prof.ev[i][1]) != cudaSuccess)
        return fail("cudaEventCreate", cudaGetLastError());
    }
    g_prof.created = true;
  }
  g_prof.on = on != 0;
  g_prof.n = 0;
  for (int i = 0; i < K_COUNT; ++i) g_prof.launches[i] = 0;
  return 0;
}

int gj_profile_read(double* ms, int64_t* timed, int64_t* launches, int n) {
  for (int i = 0; i < n && i < K_COUNT; ++i) {
    ms[i] = 0.0;
    timed[i] = 0;
    launches[i] = g_prof.launches[i];
  }
  for (int s = 0; s < g_prof.n; ++s) {
    if (cudaEventSynchronize(g_prof.ev[s][1]) != cudaSuccess) return fail("cudaEventSynchronize", cudaGetLastError());
    float t = 0.f;
    if (cudaEventElapsedTime(&t, g_prof.ev[s][0], g_prof.ev[s][1]) != cudaSuccess)
      return fail("cudaEventElapsedTime", cudaGetLastError());
    if (g_prof.id[s] < n) {
      ms[g_prof.id[s]] += t;
      timed[g_prof.id[s]]++;
    }
  }
  return K_COUNT;
}

const char* gj_profile_kernel_name(int i) { return (i >= 0 && i < K_COUNT) ? kKernelNames[i] : ""; }

int gj_profile_prepare(int64_t n, const float* shape, float* k0, void* stream) {
  if (n <= 0) return 0;
  k_profile_prepare<<<agent_grid(n), kBlock, 0, (cudaStream_t)stream>>>(n, shape, k0);
  GJ_CHECK_LAUNCH("k_profile_prepare");
  return 0;
}

int gj_transmission_forward(int64_t n, float now, const float* tinf, const float* inf, const float* maxinf,
                            const float* shape, const float* rate, const float* shift, const float* k0, float* T,
                            void* stream) {
  if (n <= 0) return 0;
  if (!tinf || !inf || !maxinf || !shape || !rate || !shift || !k0 || !T) return bad("NULL array");
  gj_step_params p;
  memset(&p, 0, sizeof(p));
  k_transmission<<<agent_grid(n), kBlock, 0, (cudaStream_t)stream>>>(n, now, tinf, inf, maxinf, shape, rate, shift, k0,
                                                                    T, p, nullptr, nullptr);
  GJ_CHECK_LAUNCH("k_transmission");
  return 0;
}

int gj_transmission_backward(int64_t n, float now, const float* tinf, const float* inf, const float* maxinf,
                             const float* shape, const float* rate, const float* shift, const float* k0,
                             const float* g_T, float* g_tinf, float* g_inf, void* stream) {
  if (n <= 0) return 0;
  if (!tinf || !inf || !maxinf || !shape || !rate || !shift || !k0 || !g_T) return bad("NULL array");
  k_transmission_bwd<<<agent_grid(n), kBlock, 0, (cudaStream_t)stream>>>(n, now, tinf, inf, maxinf, shape, rate, shift,
                                                                        k0, g_T, g_tinf, g_inf);
  GJ_CHECK_LAUNCH("k_transmission_bwd");
  return 0;
}

int gj_step_forward(const gj_world_desc* w, const gj_step_params* p, const gj_fwd_io* io, void* stream) {
  if (int e = check_world(w)) return e;
  if (!p || !io) return bad("params/io is NULL");
  if (p->n_stages > GJ_MAX_STAGES || p->n_age_bins > GJ_MAX_AGE_BINS || p->n_quar > GJ_MAX_QUAR) return bad("params sizes");
  if (!io->scratch) return bad("scratch is NULL");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t N = w->n_agents;
  if (N == 0) return 0;
  const Scratch sc = carve(w, io->scratch);
  Channels ch;
  if (int e = build_channels(w, p, &ch)) return e;
  const bool nets = (p->phases & GJ_PHASE_NETWORKS) && p->mode == GJ_MODE_STEP;
  gj_step_params pp = *p;
  if (p->mode == GJ_MODE_SEED) {
    pp.phases &= ~GJ_PHASE_NETWORKS;
    if (!io->seed_fraction) return bad("seed_fraction is NULL");
  }
  if (nets) {
    if (!io->beta || !io->S_scaled || !io->S_unscaled) return bad("beta / S buffers are NULL");
    const float* T = io->T_in;
    const float* Tq = io->T_in;
    if (!T) {  // fused: compute the transmissions from the state
      if (!io->T || !io->tinf || !io->inf || !io->maxinf || !io->k0) return bad("state / T buffers are NULL");
      float* tq = (p->n_quar > 0) ? io->Tq : io->T;
      if (!tq) return bad("Tq is NULL with an active quarantine");
      ProfScope ps(K_TRANSMISSION, st);
      k_transmission<<<agent_grid(N), kBlock, 0, st>>>(N, p->now, io->tinf, io->inf, io->maxinf, io->shape, io->rate,
                                                      io->shift, io->k0, io->T, *p, io->cur, tq);
      GJ_CHECK_LAUNCH("k_transmission");
      T = io->T;
      Tq = tq;
    } else if (p->n_quar > 0) {  // stand-alone InfectionNetworks under a quarantine: mask the given T
      if (!io->Tq || !io->cur) return bad("Tq / cur are NULL with an active quarantine");
      k_mask_transmission<<<agent_grid(N), kBlock, 0, st>>>(N, *p, io->cur, io->T_in, io->Tq);
      GJ_CHECK_LAUNCH("k_mask_transmission");
      Tq = io->Tq;
    }
    if (int e = launch_group_pass<false>(w, &pp, ch, io->beta, io->leisure_prob, T, Tq, io->S_scaled, io->S_unscaled, sc,
                                         st))
      return e;
  }
  {
    ProfScope ps(K_AGENT_FWD, st);
    k_agent_forward<<<agent_grid(N), kBlock, 0, st>>>(*w, pp, *io, sc.red_part, sc.tickets);
  }
  GJ_CHECK_LAUNCH("k_agent_forward");
  return 0;
}

int gj_step_backward(const gj_world_desc* w, const gj_step_params* p, const gj_bwd_io* io, void* stream) {
  if (int e = check_world(w)) return e;
  if (!p || !io) return bad("params/io is NULL");
  if (!io->scratch) return bad("scratch is NULL");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t N = w->n_agents;
  if (N == 0) return 0;
  const Scratch sc = carve(w, io->scratch);
  Channels ch;
  if (int e = build_channels(w, p, &ch)) return e;
  gj_step_params pp = *p;
  if (p->mode == GJ_MODE_SEED) pp.phases &= ~GJ_PHASE_NETWORKS;
  const bool nets = (pp.phases & GJ_PHASE_NETWORKS) != 0;
  if (nets && (!io->w || !io->wq || !io->R || !io->cR || !io->tape_v || !io->S_unscaled || !io->beta))
    return bad("backward workspaces are NULL");
  if ((pp.phases & GJ_PHASE_SAMPLE) && !io->tape_y0) return bad("tape_y0 is NULL");
  {
    ProfScope ps(K_AGENT_BWD, st);
    k_agent_backward<<<agent_grid(N), kBlock, 0, st>>>(*w, pp, *io, sc.red_part, sc.tickets);
  }
  GJ_CHECK_LAUNCH("k_agent_backward");
  if (nets) {
    if (int e = launch_group_pass<true>(w, &pp, ch, io->beta, io->leisure_prob, io->w, io->wq, io->cR, io->R, sc, st))
      return e;
    if (io->g_beta && pp.n_nets > 0) {
      dim3 grid(kRedBlocks / 8, pp.n_nets);
      ProfScope ps(K_DBETA, st);
      k_dbeta<<<grid, kBlock, 0, st>>>(*w, pp, io->S_unscaled, io->R, sc.dbeta_part, sc.tickets, io->g_beta);
      GJ_CHECK_LAUNCH("k_dbeta");
    }
    if (io->g_T || io->g_inf || io->g_tinf) {
      if (!io->g_T && (!io->tinf || !io->inf || !io->maxinf || !io->k0)) return bad("state arrays are NULL");
      ProfScope ps(K_AGENT_BWD_GATHER, st);
      k_agent_backward_gather<<<agent_grid(N), kBlock, 0, st>>>(*w, pp, *io);
      GJ_CHECK_LAUNCH("k_agent_backward_gather");
    }
  }
  return 0;
}

int gj_philox_fill(uint64_t seed, uint32_t call_index, int64_t n, float* E, float* u, float* z, void* stream) {
  if (n <= 0) return 0;
  k_philox_fill<<<agent_grid(n), kBlock, 0, (cudaStream_t)stream>>>(seed, call_index, n, E, u, z);
  GJ_CHECK_LAUNCH("k_philox_fill");
  return 0;
}

void gj_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1], out);
}

}  // extern "C"
