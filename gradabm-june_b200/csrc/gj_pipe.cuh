// Persistent, TMA-pipelined tile kernels of the throughput mode (included by gj_kernels.cu).
//
// Why: the per-tile CTAs of gj_fast.cuh spend most of their life on serialized latencies (tile bounds -> tables ->
// state loads -> per-agent dependent gathers -> block reduce -> fence -> ticket), ~20 us per 1024 agents.  Here
// a fixed grid of CTAs (2 per SM, 512 threads) walks the tiles round-robin with a two-stage shared-memory ring:
// while the threads work on tile i from stage s, the TMA engine (cp.async.bulk + mbarrier complete_tx) is
// already filling stage s^1 with every contiguous per-agent array of tile i+1.  Attendance tables are loaded
// once per CTA, reductions are carried in registers across tiles and combined once per CTA, and each thread
// handles its two agents of a tile in lock-step phases so that their dependent gathers overlap.
#pragma once
#include "gj_fast.cuh"

namespace gj {

constexpr int kPipeThreads = 512;
constexpr int kPipeAgentsPerThread = GJ_TILE_AGENTS / kPipeThreads;  // 2
constexpr int kPipeWords = GJ_TILE_AGENTS + 16;                       // staged elements per 4-byte array
constexpr int kPipeClsBytes = GJ_TILE_AGENTS + 48;
constexpr int kPipeMaxArrays = 16;

struct PipeTileInfo {
  uint32_t a0, a1, a0s, c0s;
};

struct PipeShared {
  alignas(16) uint64_t bar[2];
  PipeTileInfo tile[2];
  float cellv[2][GJ_MAX_CHANNELS];
  FastTables ft;
};

// dynamic shared memory: PipeShared | stage 0: arrays[kN][kPipeWords] + cls[kPipeClsBytes] | stage 1: ...
template <int kN>
struct PipeLayout {
  static constexpr size_t stage_bytes = ((size_t)kN * kPipeWords * 4 + kPipeClsBytes + 127) / 128 * 128;
  static constexpr size_t header_bytes = (sizeof(PipeShared) + 127) / 128 * 128;
  static constexpr size_t total_bytes = header_bytes + 2 * stage_bytes;
  __device__ static PipeShared& hdr(unsigned char* base) { return *reinterpret_cast<PipeShared*>(base); }
  __device__ static float* arr(unsigned char* base, int stage, int j) {
    return reinterpret_cast<float*>(base + header_bytes + stage * stage_bytes) + (size_t)j * kPipeWords;
  }
  __device__ static uint8_t* cls(unsigned char* base, int stage) {
    return base + header_bytes + stage * stage_bytes + (size_t)kN * kPipeWords * 4;
  }
};

// thread 0: start the bulk copies of one tile into `stage`
template <int kN>
__device__ __forceinline__ void pipe_issue(unsigned char* base, int stage, int64_t tile, const gj_world_desc& w,
                                           const void* const (&src)[kN], const int (&extra)[kN]) {
  PipeShared& sh = PipeLayout<kN>::hdr(base);
  const uint32_t a0 = w.tile_begin[tile], a1 = w.tile_begin[tile + 1];
  const uint32_t a0s = a0 & ~3u, a1e = (a1 + 3u) & ~3u;
  const uint32_t c0s = a0 & ~15u, c1e = (a1 + 15u) & ~15u;
  sh.tile[stage].a0 = a0;
  sh.tile[stage].a1 = a1;
  sh.tile[stage].a0s = a0s;
  sh.tile[stage].c0s = c0s;
  uint32_t bytes = 0;
#pragma unroll
  for (int j = 0; j < kN; ++j) {
    if (src[j] != nullptr) {
      const uint32_t nb = (a1e - a0s + extra[j]) * 4;
      bulk_g2s(PipeLayout<kN>::arr(base, stage, j), (const char*)src[j] + (size_t)a0s * 4, nb, &sh.bar[stage]);
      bytes += nb;
    }
  }
  bulk_g2s(PipeLayout<kN>::cls(base, stage), w.cls + c0s, c1e - c0s, &sh.bar[stage]);
  bytes += c1e - c0s;
  mbar_expect_tx(&sh.bar[stage], bytes);
}

// once per CTA: attendance tables, betas, generic network lists (no per-tile part)
__device__ __forceinline__ void pipe_static_tables(FastTables& ft, const gj_step_params& p, const Plan& pl,
                                                   const float* __restrict__ lprob, const float* __restrict__ beta) {
  for (int i = threadIdx.x; i < pl.n_lei * 200; i += blockDim.x) {
    const int j = i / 200, c = i - j * 200;
    ft.prob[j][c] = lprob[(size_t)(p.nets[pl.lei_net[j]].prob_row * 2 + p.day_type) * 200 + c];
  }
  if (threadIdx.x < p.n_nets) ft.beta[threadIdx.x] = beta ? beta[threadIdx.x] : 0.0f;
  if (threadIdx.x < GJ_MAX_TYPES) {
    int n = 0;
    for (int k = 0; k < p.n_nets; ++k)
      if (pl.tier[k] == GJ_TIER_GENERIC && p.nets[k].type == (int)threadIdx.x && n < GJ_MAX_CHANNELS)
        ft.gen_net[threadIdx.x][n++] = k;
    ft.gen_n[threadIdx.x] = n;
  }
  if (threadIdx.x == 0) ft.c_house = ft.c_plain = 0.0f;
}

// per tile: class table of the cell-tier networks from this tile's per-cell values
template <bool kSusceptibleSide>
__device__ __forceinline__ void pipe_cell_table(FastTables& ft, const float* cellv, const gj_step_params& p,
                                                const Plan& pl) {
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
  if (threadIdx.x == 255) {
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

__device__ __forceinline__ float load_cellv(const gj_world_desc& w, const gj_step_params& p, const Plan& pl,
                                            const float* __restrict__ cell_buf, int64_t tile, int j) {
  const int t = p.nets[pl.t2_net[j]].type;
  return cell_buf[(w.cell_off[t] + w.tile_cell[t][tile]) * GJ_MAX_CHANNELS + j];
}

// block-wide sums for kPipeThreads threads, result to out[0..nr)
template <typename T, int kR>
__device__ __forceinline__ void pipe_block_sums(T (&v)[kR], int nr, T* __restrict__ out) {
  __shared__ T sm[kPipeThreads / 32][kR];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int r = 0; r < kR; ++r)
    if (r < nr) v[r] = warp_sum(v[r]);
  if (lane == 0) {
#pragma unroll
    for (int r = 0; r < kR; ++r)
      if (r < nr) sm[wid][r] = v[r];
  }
  __syncthreads();
  if (threadIdx.x < nr) {
    T s = (T)0;
#pragma unroll
    for (int k = 0; k < kPipeThreads / 32; ++k) s += sm[k][threadIdx.x];
    out[threadIdx.x] = s;
  }
  __syncthreads();
}

// value of staged array element for agent index b if it lies in the staged window, else from global memory
__device__ __forceinline__ float staged_or_global(const float* st, uint32_t a0s, uint32_t n_staged, const float* g,
                                                  uint32_t b) {
  const uint32_t i = b - a0s;
  return (i < n_staged) ? st[i] : g[b];
}

// =====================================================================================================
// forward
// =====================================================================================================
enum { FS_S = 0, FS_INF, FS_TINF, FS_CUR, FS_NXT, FS_TTN, FS_PTR, FS_SLOT, FS_RPC, FS_TR, FS_COUNT };

__global__ void __launch_bounds__(kPipeThreads, 2) k_pipe_forward(gj_world_desc w, gj_step_params p, Plan pl,
                                                                  gj_fwd_io io, const float* __restrict__ cell_buf,
                                                                  double* __restrict__ red_part,
                                                                  unsigned int* __restrict__ ticket) {
  extern __shared__ __align__(128) unsigned char smem[];
  using L = PipeLayout<FS_COUNT>;
  PipeShared& sh = L::hdr(smem);
  const int64_t N = w.n_agents;
  const float* __restrict__ Tsrc = io.T_in ? io.T_in : io.T;
  const float* __restrict__ Tq = (p.n_quar > 0) ? io.Tq : Tsrc;
  // the first range-tier network gets its slot words, contact probabilities and member values staged
  const int k_r0 = pl.n_t1 > 0 ? pl.t1_net[0] : -1;
  const int kind_r0 = k_r0 >= 0 ? p.nets[k_r0].kind : 0;
  const float* __restrict__ Tr0 = (kind_r0 == GJ_KIND_HOUSEHOLD) ? Tsrc : Tq;
  const void* const src[FS_COUNT] = {io.s, io.inf, io.tinf, io.cur, io.nxt, io.ttn,
                                     pl.n_generic > 0 ? (const void*)w.am_ptr : nullptr,
                                     k_r0 >= 0 ? (const void*)pl.slot[k_r0] : nullptr,
                                     k_r0 >= 0 ? (const void*)pl.rpc[k_r0] : nullptr,
                                     k_r0 >= 0 ? (const void*)Tr0 : nullptr};
  const int extra[FS_COUNT] = {0, 0, 0, 0, 0, 0, 4, 0, 0, 0};
  const bool has_cell = pl.n_t2 > 0;

  if (threadIdx.x == 0) {
    mbar_init(&sh.bar[0], 1);
    mbar_init(&sh.bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  int64_t tile = blockIdx.x;
  if (tile < w.n_tiles) {
    if (threadIdx.x == 0) pipe_issue<FS_COUNT>(smem, 0, tile, w, src, extra);
    if (has_cell && threadIdx.x < pl.n_t2) sh.cellv[0][threadIdx.x] = load_cellv(w, p, pl, cell_buf, tile, threadIdx.x);
  }
  pipe_static_tables(sh.ft, p, pl, io.leisure_prob, io.beta);
  float red[kMaxRed];
#pragma unroll
  for (int r = 0; r < kMaxRed; ++r) red[r] = 0.0f;
  __syncthreads();

  for (int it = 0; tile < w.n_tiles; ++it, tile += gridDim.x) {
    const int stage = it & 1;
    const int64_t next = tile + gridDim.x;
    float next_cellv = 0.0f;
    if (next < w.n_tiles) {  // stage^1 was released by the __syncthreads that ended the previous iteration
      if (threadIdx.x == 0) pipe_issue<FS_COUNT>(smem, stage ^ 1, next, w, src, extra);
      if (has_cell && threadIdx.x < pl.n_t2) next_cellv = load_cellv(w, p, pl, cell_buf, next, threadIdx.x);
    }
    if (has_cell) pipe_cell_table<true>(sh.ft, sh.cellv[stage], p, pl);
    mbar_wait(&sh.bar[stage], (it >> 1) & 1);
    __syncthreads();  // class table ready, tile info visible
    const PipeTileInfo ti = sh.tile[stage];
    const uint32_t n_staged = ((ti.a1 + 3u) & ~3u) - ti.a0s;
    const float* s_s = L::arr(smem, stage, FS_S);
    const float* s_inf = L::arr(smem, stage, FS_INF);
    const float* s_tinf = L::arr(smem, stage, FS_TINF);
    const float* s_cur = L::arr(smem, stage, FS_CUR);
    const float* s_nxt = L::arr(smem, stage, FS_NXT);
    const float* s_ttn = L::arr(smem, stage, FS_TTN);
    const uint32_t* s_ptr = reinterpret_cast<const uint32_t*>(L::arr(smem, stage, FS_PTR));
    const uint32_t* s_slot = reinterpret_cast<const uint32_t*>(L::arr(smem, stage, FS_SLOT));
    const float* s_rpc = L::arr(smem, stage, FS_RPC);
    const float* s_tr = L::arr(smem, stage, FS_TR);
    const uint8_t* s_cls = L::cls(smem, stage);

    // ---- phase A: per-agent words from shared memory, first generic entry from global ----------------
    uint32_t a[kPipeAgentsPerThread], e0[kPipeAgentsPerThread], deg[kPipeAgentsPerThread], ent0[kPipeAgentsPerThread];
    int cls[kPipeAgentsPerThread];
    bool live[kPipeAgentsPerThread];
#pragma unroll
    for (int h = 0; h < kPipeAgentsPerThread; ++h) {
      a[h] = ti.a0 + threadIdx.x + h * kPipeThreads;
      live[h] = a[h] < ti.a1;
      const uint32_t i = a[h] - ti.a0s;
      cls[h] = live[h] ? s_cls[a[h] - ti.c0s] : 0;
      e0[h] = 0;
      deg[h] = 0;
      ent0[h] = 0;
      if (live[h] && pl.n_generic > 0) {
        e0[h] = s_ptr[i];
        deg[h] = s_ptr[i + 1] - e0[h];
        if (deg[h] > 0) ent0[h] = w.am_ent[e0[h]];
      }
    }
    // ---- phase B: gathers (group sums from L2, household neighbours from the staged tile) ---------------
    GenericSums g[kPipeAgentsPerThread];
    RangeSums rs[kPipeAgentsPerThread];
#pragma unroll
    for (int h = 0; h < kPipeAgentsPerThread; ++h) {
      g[h].house = g[h].plain = 0.0f;
      rs[h].house = rs[h].plain = 0.0f;
      if (!live[h]) continue;
      if (deg[h] > 0) add_entry(g[h], sh.ft, p, pl, io.S_scaled, ent0[h], cls[h], true);
      for (uint32_t j = 1; j < deg[h]; ++j) add_entry(g[h], sh.ft, p, pl, io.S_scaled, w.am_ent[e0[h] + j], cls[h], true);
      if (k_r0 >= 0) {
        const uint32_t i = a[h] - ti.a0s;
        const uint32_t slot = s_slot[i];
        if (slot != kNoSlot) {
          const uint32_t b0 = a[h] - (slot >> 16), nb = slot & 0xFFFFu;
          const float cg = sh.ft.beta[k_r0] * s_rpc[i];
          float S = 0.0f;
          if (kind_r0 <= GJ_KIND_HOUSEHOLD) {
            for (uint32_t b = b0; b < b0 + nb; ++b) S += staged_or_global(s_tr, ti.a0s, n_staged, Tr0, b);
          } else {
            for (uint32_t b = b0; b < b0 + nb; ++b)
              S += sh.ft.prob[pl.net_lei[k_r0]][w.cls[b]] * staged_or_global(s_tr, ti.a0s, n_staged, Tr0, b);
          }
          float own = 1.0f;
          if (kind_r0 >= GJ_KIND_LEISURE) {
            own = sh.ft.prob[pl.net_lei[k_r0]][cls[h]];
            if (kind_r0 == GJ_KIND_CARE_VISIT) own = own * (((cls[h] % 100) > 75) ? 1.0f : 0.0f);
          }
          if (kind_r0 == GJ_KIND_HOUSEHOLD) rs[h].house += cg * S;
          else rs[h].plain += (cg * S) * own;
        }
      }
      if (pl.n_t1 > 1) {  // further range-tier networks (rare): plain global path
        const RangeSums more = range_sums<false>(w, sh.ft, p, pl, Tsrc, Tq, a[h], cls[h], 1);
        rs[h].house += more.house;
        rs[h].plain += more.plain;
      }
    }
    // ---- phase C: pressure -> q -> draw -> update -> symptoms -> outputs --------------------------------
#pragma unroll
    for (int h = 0; h < kPipeAgentsPerThread; ++h) {
      if (!live[h]) continue;
      const uint32_t i = a[h] - ti.a0s;
      AgentState st;
      st.s = s_s[i];
      st.inf = io.inf ? s_inf[i] : 0.0f;
      st.tinf = io.tinf ? s_tinf[i] : 0.0f;
      st.cur = io.cur ? s_cur[i] : 1.0f;
      st.nxt = io.nxt ? s_nxt[i] : 1.0f;
      st.ttn = io.ttn ? s_ttn[i] : 0.0f;
      const float mq = (p.n_quar > 0) ? quarantine_mask(p, st.cur) : 1.0f;
      float house = g[h].house + rs[h].house, plain = g[h].plain + rs[h].plain;
      if (has_cell) {
        house += sh.ft.c_house;
        plain += sh.ft.c_plain + sh.ft.L[cls[h]];
      }
      const float X = house + mq * plain;  // pressure per unit susceptibility
      const float lam = X * st.s;
      const float q = not_infected_prob(lam, p.dt);
      io.tape_v[a[h]] = (st.s == 0.0f) ? X : lam;
      if (io.q) io.q[a[h]] = q;
      if (io.lam) io.lam[a[h]] = lam;
      forward_tail<true>(p, io, N, a[h], cls[h] % 100, q, st, red);
    }
    if (has_cell && next < w.n_tiles && threadIdx.x < pl.n_t2) sh.cellv[stage ^ 1][threadIdx.x] = next_cellv;
    __syncthreads();  // everybody is done with this stage (and with the class table) before it is refilled
  }
  if (io.red) {
    double redd[kMaxRed];
#pragma unroll
    for (int r = 0; r < kMaxRed; ++r) redd[r] = (double)red[r];
    pipe_block_sums<double, kMaxRed>(redd, 2 + p.n_age_bins, red_part + (int64_t)blockIdx.x * kMaxRed);
    finish_partials<kMaxRed>(2 + p.n_age_bins, red_part, gridDim.x, ticket, io.red);
  }
}

}  // namespace gj
