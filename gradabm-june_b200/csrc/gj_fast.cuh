// Throughput-mode tile kernels (included by gj_kernels.cu after gj_tiled.cuh).
//
// Same arithmetic as the reference-order kernels of gj_tiled.cuh, re-associated so that the per-agent
// instruction count is small: the eleven-network loop is replaced by three terms
//   GENERIC  sum over the agent's (usually <= 1) CSR entries, networks looked up per edge type,
//   RANGE    the household re-sum over neighbouring agents,
//   CELL     ONE shared-memory lookup L[class] = sum_k B_k * p_k(class) built once per tile, which covers all
//            leisure networks at once (their masks depend on the agent only through its (sex, age) class),
// so pressure = s * (c_h + range + mq * (generic + c_p + L[class])).  The result differs from the
// reference-order kernels only by fp32 re-association (~1e-7 relative); these kernels are used when the noise
// is the in-kernel Philox stream, the reference-order ones when noise is injected (parity tests).
#pragma once
#include "gj_tiled.cuh"

namespace gj {

// ---- TMA bulk staging of a tile's contiguous per-agent arrays into shared memory ------------------------
// cp.async.bulk (1-D, no tensor map) moves [a0, a1) of every state array into shared memory while the CTA's
// threads are idle; with several CTAs per SM the copies of one tile overlap the arithmetic of another, and no
// register is spent on loads in flight.  Arrays whose base pointer is not 16-byte aligned (views) fall back to
// cooperative loads.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}

constexpr int kStageElems = GJ_TILE_AGENTS + 8;  // room for the 16-byte alignment slack on both sides

// Stage src[a0s .. a1e) (element size 4) into dst; a0s = a0 & ~3, a1e = (a1 + 3) & ~3.  Thread 0 issues the bulk
// copy when the source is 16-byte aligned, otherwise all threads copy.  Returns the bytes the barrier must expect.
__device__ __forceinline__ uint32_t stage4(void* dst, const void* src, uint32_t a0s, uint32_t a1e, uint64_t* bar) {
  if (src == nullptr) return 0;
  const char* g = (const char*)src + (size_t)a0s * 4;
  const uint32_t bytes = (a1e - a0s) * 4;
  if ((((uintptr_t)src) & 15) == 0) {
    if (threadIdx.x == 0) bulk_g2s(dst, g, bytes, bar);
    return bytes;
  }
  for (uint32_t i = threadIdx.x; i < a1e - a0s; i += blockDim.x) ((uint32_t*)dst)[i] = ((const uint32_t*)g)[i];
  return 0;
}

struct FastTables {
  float L[200];        // CELL tier, attendance-table kinds: sum_k V_k * p_k(class) [* (age>75) for care visits on the S side]
  float c_house;       // CELL tier, HOUSEHOLD-kind networks: sum_k V_k
  float c_plain;       // CELL tier, PLAIN-kind networks
  float prob[GJ_MAX_CHANNELS][200];
  float beta[GJ_MAX_NETS];
  // GENERIC tier: networks per edge type
  int gen_n[GJ_MAX_TYPES];
  int gen_net[GJ_MAX_TYPES][GJ_MAX_CHANNELS];
};

// V_k = per-cell value of cell channel k (forward: sum of scaled group sums; backward: sum of c_g * R_g)
template <bool kSusceptibleSide>
__device__ __forceinline__ void build_fast_tables(FastTables& ft, const gj_world_desc& w, const gj_step_params& p,
                                                  const Plan& pl, const float* __restrict__ lprob,
                                                  const float* __restrict__ beta, const float* __restrict__ cell_buf,
                                                  int64_t tile) {
  __shared__ float cellv[GJ_MAX_CHANNELS];
  for (int i = threadIdx.x; i < pl.n_lei * 200; i += blockDim.x) {
    const int j = i / 200, c = i - j * 200;
    ft.prob[j][c] = lprob[(size_t)(p.nets[pl.lei_net[j]].prob_row * 2 + p.day_type) * 200 + c];
  }
  if (threadIdx.x < p.n_nets) ft.beta[threadIdx.x] = beta ? beta[threadIdx.x] : 0.0f;
  if (threadIdx.x < pl.n_t2 && cell_buf) {
    const int t = p.nets[pl.t2_net[threadIdx.x]].type;
    cellv[threadIdx.x] = cell_buf[(w.cell_off[t] + w.tile_cell[t][tile]) * GJ_MAX_CHANNELS + threadIdx.x];
  }
  if (threadIdx.x < GJ_MAX_TYPES) {
    int n = 0;
    for (int k = 0; k < p.n_nets; ++k)
      if (pl.tier[k] == GJ_TIER_GENERIC && p.nets[k].type == (int)threadIdx.x && n < GJ_MAX_CHANNELS)
        ft.gen_net[threadIdx.x][n++] = k;
    ft.gen_n[threadIdx.x] = n;
  }
  __syncthreads();
  if (cell_buf) {
    for (int c = threadIdx.x; c < 200; c += blockDim.x) {
      float acc = 0.0f;
      for (int j = 0; j < pl.n_t2; ++j) {
        const int k = pl.t2_net[j];
        const int kind = p.nets[k].kind;
        if (kind >= GJ_KIND_LEISURE) {
          float v = cellv[j] * ft.prob[pl.net_lei[k]][c];
          if (kSusceptibleSide && kind == GJ_KIND_CARE_VISIT) v = v * (((c % 100) > 75) ? 1.0f : 0.0f);
          acc += v;
        }
      }
      ft.L[c] = acc;
    }
    if (threadIdx.x == 0) {
      float ch = 0.0f, cp = 0.0f;
      for (int j = 0; j < pl.n_t2; ++j) {
        const int kind = p.nets[pl.t2_net[j]].kind;
        if (kind == GJ_KIND_HOUSEHOLD) ch += cellv[j];
        else if (kind == GJ_KIND_PLAIN) cp += cellv[j];
      }
      ft.c_house = ch;
      ft.c_plain = cp;
    }
  }
  __syncthreads();
}

// generic-tier sum over the agent's CSR entries of `buf[s_off_k + group]`, split by the network's mask kind:
// returns (sum over HOUSEHOLD-kind networks, sum over PLAIN, sum over table kinds weighted by p_k(class))
struct GenericSums {
  float house, plain;
};

__device__ __forceinline__ void add_entry(GenericSums& g, const FastTables& ft, const gj_step_params& p, const Plan& pl,
                                          const float* __restrict__ buf, uint32_t ent, int cls, bool care_age_side) {
  const int type = (int)(ent >> 28);
  const uint32_t grp = ent & 0x0FFFFFFFu;
  const int n = ft.gen_n[type];
  for (int c = 0; c < n; ++c) {
    const int k = ft.gen_net[type][c];
    const int kind = p.nets[k].kind;
    float v = buf[(int64_t)p.nets[k].s_off + grp];
    if (kind == GJ_KIND_HOUSEHOLD) {
      g.house += v;
    } else {
      if (kind >= GJ_KIND_LEISURE) {
        v = v * ft.prob[pl.net_lei[k]][cls];
        if (care_age_side && kind == GJ_KIND_CARE_VISIT) v = v * (((cls % 100) > 75) ? 1.0f : 0.0f);
      }
      g.plain += v;
    }
  }
}

__device__ __forceinline__ GenericSums generic_sums(const gj_world_desc& w, const FastTables& ft, const gj_step_params& p,
                                                    const Plan& pl, const float* __restrict__ buf, uint32_t a, int cls,
                                                    bool care_age_side) {
  GenericSums g;
  g.house = g.plain = 0.0f;
  if (pl.n_generic > 0) {
    const uint32_t e0 = w.am_ptr[a], e1 = w.am_ptr[a + 1];
    for (uint32_t j = e0; j < e1; ++j) add_entry(g, ft, p, pl, buf, w.am_ent[j], cls, care_age_side);
  }
  return g;
}

// range-tier networks: re-sum the agent's (small) group from its neighbours.  in0/in1 = unmasked / quarantine-masked
// per-agent values (T, Tq forward; w, wq backward)
struct RangeSums {
  float house, plain;  // already multiplied by beta_k * pc_g
};

template <bool kBwd>
__device__ __forceinline__ void range_one(RangeSums& r, int k, const gj_world_desc& w, const FastTables& ft,
                                          const gj_step_params& p, const Plan& pl, const float* __restrict__ in0,
                                          const float* __restrict__ in1, uint32_t a, int cls) {
  const uint32_t slot = pl.slot[k][a];
  if (slot == kNoSlot) return;
  const int kind = p.nets[k].kind;
  const uint32_t b0 = a - (slot >> 16), nb = slot & 0xFFFFu;
  const float cg = ft.beta[k] * pl.rpc[k][a];
  float S = 0.0f;
  if (kind == GJ_KIND_HOUSEHOLD) {
    for (uint32_t b = b0; b < b0 + nb; ++b) S += in0[b];
  } else if (kind == GJ_KIND_PLAIN) {
    for (uint32_t b = b0; b < b0 + nb; ++b) S += in1[b];
  } else {
    for (uint32_t b = b0; b < b0 + nb; ++b) {
      const int cb = w.cls[b];
      float v = ft.prob[pl.net_lei[k]][cb] * in1[b];
      if (kBwd && kind == GJ_KIND_CARE_VISIT) v = v * (((cb % 100) > 75) ? 1.0f : 0.0f);
      S += v;
    }
  }
  // the agent-side mask of a table kind is applied here so that the caller only distinguishes house / plain
  float own = 1.0f;
  if (kind >= GJ_KIND_LEISURE) {
    own = ft.prob[pl.net_lei[k]][cls];
    if (!kBwd && kind == GJ_KIND_CARE_VISIT) own = own * (((cls % 100) > 75) ? 1.0f : 0.0f);
  }
  if (kind == GJ_KIND_HOUSEHOLD) r.house += cg * S;
  else r.plain += (cg * S) * own;
}

template <bool kBwd>
__device__ __forceinline__ RangeSums range_sums(const gj_world_desc& w, const FastTables& ft, const gj_step_params& p,
                                                const Plan& pl, const float* __restrict__ in0,
                                                const float* __restrict__ in1, uint32_t a, int cls, int first = 0) {
  RangeSums r;
  r.house = r.plain = 0.0f;
  for (int i = first; i < pl.n_t1; ++i) range_one<kBwd>(r, pl.t1_net[i], w, ft, p, pl, in0, in1, a, cls);
  return r;
}

// =====================================================================================================
// forward
// =====================================================================================================
struct FwdStage {
  float s[kStageElems], inf[kStageElems], tinf[kStageElems], cur[kStageElems], nxt[kStageElems], ttn[kStageElems];
  uint32_t ptr[kStageElems];  // am_ptr[a0s .. a1e] (one extra element is read directly)
  alignas(16) uint64_t bar;
};

__global__ void __launch_bounds__(kBlock, 4) k_fast_forward(gj_world_desc w, gj_step_params p, Plan pl, gj_fwd_io io,
                                                            const float* __restrict__ cell_buf,
                                                            double* __restrict__ red_part,
                                                            unsigned int* __restrict__ ticket) {
  __shared__ FastTables ft;
  __shared__ alignas(128) FwdStage sg;
  const int64_t N = w.n_agents;
  const int64_t tile = blockIdx.x;
  const uint32_t a0 = w.tile_begin[tile], a1 = w.tile_begin[tile + 1];
  const uint32_t a0s = a0 & ~3u;
  uint32_t a1e = (a1 + 3u) & ~3u;
  if (threadIdx.x == 0) {
    mbar_init(&sg.bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  {
    uint32_t bytes = 0;
    bytes += stage4(sg.s, io.s, a0s, a1e, &sg.bar);
    bytes += stage4(sg.inf, io.inf, a0s, a1e, &sg.bar);
    bytes += stage4(sg.tinf, io.tinf, a0s, a1e, &sg.bar);
    bytes += stage4(sg.cur, io.cur, a0s, a1e, &sg.bar);
    bytes += stage4(sg.nxt, io.nxt, a0s, a1e, &sg.bar);
    bytes += stage4(sg.ttn, io.ttn, a0s, a1e, &sg.bar);
    if (pl.n_generic > 0) bytes += stage4(sg.ptr, w.am_ptr, a0s, a1e, &sg.bar);
    if (threadIdx.x == 0) mbar_expect_tx(&sg.bar, bytes);
  }
  build_fast_tables<true>(ft, w, p, pl, io.leisure_prob, io.beta, pl.n_t2 > 0 ? cell_buf : nullptr, tile);
  float red[kMaxRed];
#pragma unroll
  for (int r = 0; r < kMaxRed; ++r) red[r] = 0.0f;
  const float* __restrict__ Tsrc = io.T_in ? io.T_in : io.T;
  const float* __restrict__ Tq = (p.n_quar > 0) ? io.Tq : Tsrc;
  const bool has_cell = pl.n_t2 > 0;
  mbar_wait(&sg.bar, 0);
  __syncthreads();  // cooperative (unaligned) copies, if any, are visible too

  for (uint32_t a = a0 + threadIdx.x; a < a1; a += kBlock) {
    const uint32_t i = a - a0s;
    const int cls = w.cls[a];
    AgentState st;
    st.s = sg.s[i];
    st.inf = io.inf ? sg.inf[i] : 0.0f;
    st.tinf = io.tinf ? sg.tinf[i] : 0.0f;
    st.cur = io.cur ? sg.cur[i] : 1.0f;
    st.nxt = io.nxt ? sg.nxt[i] : 1.0f;
    st.ttn = io.ttn ? sg.ttn[i] : 0.0f;
    const float mq = (p.n_quar > 0) ? quarantine_mask(p, st.cur) : 1.0f;
    GenericSums g;
    g.house = g.plain = 0.0f;
    if (pl.n_generic > 0) {
      const uint32_t e0 = sg.ptr[i];
      const uint32_t e1 = (i + 1 < a1e - a0s) ? sg.ptr[i + 1] : w.am_ptr[a + 1];
      for (uint32_t j = e0; j < e1; ++j) add_entry(g, ft, p, pl, io.S_scaled, w.am_ent[j], cls, true);
    }
    const RangeSums r = range_sums<false>(w, ft, p, pl, Tsrc, Tq, a, cls);
    float house = g.house + r.house, plain = g.plain + r.plain;
    if (has_cell) {
      house += ft.c_house;
      plain += ft.c_plain + ft.L[cls];
    }
    const float X = house + mq * plain;  // pressure per unit susceptibility
    const float lam = X * st.s;
    const float q = not_infected_prob(lam, p.dt);
    io.tape_v[a] = (st.s == 0.0f) ? X : lam;
    if (io.q) io.q[a] = q;
    if (io.lam) io.lam[a] = lam;
    forward_tail<true>(p, io, N, a, cls % 100, q, st, red);
  }
  if (io.red) {
    double redd[kMaxRed];
#pragma unroll
    for (int r = 0; r < kMaxRed; ++r) redd[r] = (double)red[r];
    block_sums<double, kMaxRed>(redd, 2 + p.n_age_bins, red_part + tile * kMaxRed);
    finish_partials<kMaxRed>(2 + p.n_age_bins, red_part, gridDim.x, ticket, io.red);
  }
}

// =====================================================================================================
// backward part 1: per-agent cotangents, w / wq, tile partial sums of the cell channels
// =====================================================================================================
__global__ void __launch_bounds__(kBlock, 4) k_fast_backward(gj_world_desc w, gj_step_params p, Plan pl, gj_bwd_io io,
                                                             float* __restrict__ tile_part) {
  __shared__ FastTables ft;
  const int64_t N = w.n_agents;
  const int64_t tile = blockIdx.x;
  const uint32_t a0 = w.tile_begin[tile], a1 = w.tile_begin[tile + 1];
  build_fast_tables<true>(ft, w, p, pl, io.leisure_prob, nullptr, nullptr, tile);
  float acc[GJ_MAX_CHANNELS];
#pragma unroll
  for (int j = 0; j < GJ_MAX_CHANNELS; ++j) acc[j] = 0.0f;
  for (uint32_t a = a0 + threadIdx.x; a < a1; a += kBlock) {
    const int cls = w.cls[a];
    AgentState st;
    st.s = io.s[a];
    st.inf = io.inf ? io.inf[a] : 0.0f;
    st.tinf = io.tinf ? io.tinf[a] : 0.0f;
    st.cur = io.cur ? io.cur[a] : 1.0f;
    st.nxt = io.nxt ? io.nxt[a] : 1.0f;
    st.ttn = io.ttn ? io.ttn[a] : 0.0f;
    const BackAgent r = backward_agent(p, io, N, a, cls % 100, st, true);
    const float mq = (p.n_quar > 0) ? quarantine_mask(p, st.cur) : 1.0f;
    const float wv = r.glam * st.s;
    const float wqv = r.glam * (mq * st.s);
    io.w[a] = wv;
    if (io.wq != io.w) io.wq[a] = wqv;
    if (pl.n_t2 > 0 && r.glam != 0.0f) {
      const float agef = ((cls % 100) > 75) ? 1.0f : 0.0f;
#pragma unroll
      for (int j = 0; j < GJ_MAX_CHANNELS; ++j) {
        if (j < pl.n_t2) {
          const int k = pl.t2_net[j];
          const int kind = p.nets[k].kind;
          float v = (kind == GJ_KIND_HOUSEHOLD) ? wv : wqv;
          if (kind >= GJ_KIND_LEISURE) v = ft.prob[pl.net_lei[k]][cls] * wqv;
          if (kind == GJ_KIND_CARE_VISIT) v = v * agef;
          acc[j] += v;
        }
      }
    }
    if (io.g_s) io.g_s[a] = r.gs;
    if (io.g_inf) io.g_inf[a] = r.ginf;
    if (io.g_tinf) io.g_tinf[a] = r.gtinf;
    if (io.g_cur) io.g_cur[a] = r.gcur;
    if (io.g_nxt) io.g_nxt[a] = r.gnxt;
    if (io.g_ttn) io.g_ttn[a] = r.gttn;
  }
  if (pl.n_t2 > 0) block_sums<float, GJ_MAX_CHANNELS>(acc, pl.n_t2, tile_part + tile * GJ_MAX_CHANNELS);
}

// =====================================================================================================
// backward part 2: dL/dT -> (is_infected, infection_time); d/dbeta of the range-tier networks
// =====================================================================================================
__global__ void __launch_bounds__(kBlock, 4) k_fast_backward_gather(gj_world_desc w, gj_step_params p, Plan pl,
                                                                    gj_bwd_io io, const float* __restrict__ cell_buf,
                                                                    double* __restrict__ dbeta_tile) {
  __shared__ FastTables ft;
  const int64_t tile = blockIdx.x;
  const uint32_t a0 = w.tile_begin[tile], a1 = w.tile_begin[tile + 1];
  build_fast_tables<false>(ft, w, p, pl, io.leisure_prob, io.beta, pl.n_t2 > 0 ? cell_buf : nullptr, tile);
  double db[GJ_MAX_RANGE_NETS];
#pragma unroll
  for (int i = 0; i < GJ_MAX_RANGE_NETS; ++i) db[i] = 0.0;
  const bool quar = p.n_quar > 0;
  const bool has_cell = pl.n_t2 > 0;
  const float* __restrict__ T = io.T_in;  // the forward's transmissions

  for (uint32_t a = a0 + threadIdx.x; a < a1; a += kBlock) {
    const int cls = w.cls[a];
    const float mq = quar ? quarantine_mask(p, io.cur[a]) : 1.0f;
    const GenericSums g = generic_sums(w, ft, p, pl, io.cR, a, cls, false);
    const RangeSums r = range_sums<true>(w, ft, p, pl, io.w, io.wq, a, cls);
    float house = g.house + r.house, plain = g.plain + r.plain;
    if (has_cell) {
      house += ft.c_house;
      plain += ft.c_plain + ft.L[cls];
    }
    const float gT = house + mq * plain;
    // d/dbeta of the range-tier networks: the group's first member adds pc_g * (sum of transmissions) * R_g
#pragma unroll
    for (int i = 0; i < GJ_MAX_RANGE_NETS; ++i) {
      if (i < pl.n_t1) {
        const int k = pl.t1_net[i];
        const uint32_t slot = pl.slot[k][a];
        if (slot != kNoSlot && (slot >> 16) == 0) {
          const int kind = p.nets[k].kind;
          const uint32_t nb = slot & 0xFFFFu;
          float R = 0.0f, S = 0.0f;
          for (uint32_t b = a; b < a + nb; ++b) {
            const int cb = (kind >= GJ_KIND_LEISURE) ? w.cls[b] : 0;
            const float Tb = T[b];
            const float Tqb = quar ? quarantine_mask(p, io.cur[b]) * Tb : Tb;
            float vw = (kind == GJ_KIND_HOUSEHOLD) ? io.w[b] : io.wq[b];
            float vt = (kind == GJ_KIND_HOUSEHOLD) ? Tb : Tqb;
            if (kind >= GJ_KIND_LEISURE) {
              const float pb = ft.prob[pl.net_lei[k]][cb];
              vw = vw * pb;
              vt = vt * pb;
              if (kind == GJ_KIND_CARE_VISIT) vw = vw * (((cb % 100) > 75) ? 1.0f : 0.0f);
            }
            R += vw;
            S += vt;
          }
          db[i] += (double)(pl.rpc[k][a] * S) * (double)R;
        }
      }
    }
    if (io.g_T) {
      io.g_T[a] = gT;
    } else if (gT != 0.0f) {
      const TransTerms tt =
          transmission_terms<true>(p.now, io.tinf[a], io.maxinf[a], io.shape[a], io.rate[a], io.shift[a], io.k0[a]);
      io.g_inf[a] += gT * tt.coef;
      io.g_tinf[a] += gT * (tt.dcoef * io.inf[a]);
    }
  }
  if (pl.n_t1 > 0) block_sums<double, GJ_MAX_RANGE_NETS>(db, pl.n_t1, dbeta_tile + tile * GJ_MAX_RANGE_NETS);
}

}  // namespace gj
