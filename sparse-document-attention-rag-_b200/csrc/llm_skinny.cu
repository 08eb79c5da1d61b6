// Few-row linear layers of a decode step (1..32 prompts) on the tensor cores, as a pure weight stream:
//
//     out[b, n] = sum_k W[n, k] * f(x)[b, k]   (+ bias) (+ residual)        W: [N, K] bf16 (nn.Linear layout)
//
// The op is HBM-bound (every weight byte is used once; 16 GB per token for the Llama-3-8B geometry), so the kernel is
// built around the stream: a CTA owns up to 128 weight ROWS (x a K slice) and pulls them through a deep TMA ring (one
// multi-k-block box per stage, ~150-200 KB in flight per SM, no registers, no L1); `tcgen05.mma` consumes the tiles straight
// from shared memory with the WEIGHTS on the M side — D[128 rows x 16] = W_tile[128 x 64] * X[16 x 64]^T per k-block, the
// prompts padded to N = 16 (32 beyond 16 prompts).  In SS mode an MMA costs ~50 cycles + N/2 whatever it computes (the
// 128 x 16 A slice is read from shared memory), so the weights have to be the operand that fills the 128 rows: 4 KB of
// weights per ~56 cycles, twice what the stream needs; with the prompts on the M side (first version) a 32-row weight tile
// cost the same 56 cycles for 1 KB and the tensor pipe, not HBM, set the pace.
//
// Layers with few rows (o_proj / down_proj: 4096 rows = 28 per SM) would leave the 128 rows mostly empty: they are split
// along K over a thread-block CLUSTER (2-8 CTAs, one row tile, 1/SK of K each); the partial accumulators travel to the
// cluster's rank 0 through distributed shared memory and are summed there in rank order (deterministic).
// SwiGLU pairs: a CTA's tile is R gate rows + the R matching up rows (one 4-D box), accumulated as two M tiles into
// neighbouring TMEM columns, so TMEM lane j holds gate j and up j of every prompt.
//
// Weights do not depend on the previous kernel of the step: the ring is filled BEFORE `griddepcontrol.wait`, so with
// programmatic dependent launch the stream of layer i+1 starts while layer i drains.  Prologue forms:
//   * norm = 1: RMSNorm of the B input rows (LlamaRMSNorm arithmetic: fp32 statistics, bf16(x * rstd) * weight) computed by
//     the CTA (a warp per row) and staged K-major / 128B-swizzled for its K slice;
//   * norm = 0: the x tiles travel through the ring beside their weight tiles (1 KB per 64 k, out of L2).
#include "common.cuh"
#include "sm100_ptx.cuh"

using namespace sm100;

namespace {

constexpr int SK_THREADS = 256;  // warp 0: TMA producer, warp 1: MMA issuer, warps 2-7: norm prologue, warps 4-7: epilogue
constexpr int SK_NORM_WARPS = 6;
constexpr int SK_MAX_STAGES = 32;
constexpr int SK_A_SPAN = 128 * 128;  // bytes the M = 128 A descriptor reads from a weight tile's base
constexpr int SK_MAX_CLUSTER = 8;

struct SkinnyParams {
  const __nv_bfloat16* x;
  long long ldx;
  const __nv_bfloat16* norm_w;
  float eps;
  const __nv_bfloat16* bias;
  const __nv_bfloat16* resid;
  long long ldr;
  __nv_bfloat16* out;
  long long ldo;
  int B, RB, NC, K, N, kblocks;
  int R, parts, half, units, n_tiles, SK, KS, stages, norm;
  uint32_t w_tile_bytes, x_tile_bytes, stage_bytes, ring_bytes, xs_bytes, recv_bytes, tmem_cols;
  int debug_no_mma;  // bring-up: stream only (SDAG_SKINNY_NOMMA=1)
};

__device__ __forceinline__ float bf_round(float v) { return __bfloat162float(__float2bfloat16(v)); }
__device__ __forceinline__ void unpack8(const uint4& a, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&a);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x, f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 a;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&a);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return a;
}
__device__ __forceinline__ void st_cluster_f32(uint32_t cluster_addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(cluster_addr), "f"(v) : "memory");
}
__device__ __forceinline__ void mbar_arrive_release_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_acquire_cluster(uint64_t* bar, uint32_t parity) {
  const long long t0 = clock64();
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (!ok && clock64() - t0 > SDAG_SPIN_TRAP_CYCLES) __trap();
  }
}

__global__ void __launch_bounds__(SK_THREADS, 2)
skinny_linear_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmX, const SkinnyParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // the M = 128 A descriptor reads 16 KB from a weight tile's base whatever R is: the ring comes first, so the over-read of
  // its last tiles lands in the areas behind it (plus a pad sized by the host)
  uint8_t* ring = smem;                                                      // [stages][KS x parts weight tiles | KS x tiles (norm = 0)]
  uint8_t* xs = smem + p.ring_bytes;                                         // norm = 1: [slice k-blocks][RB rows x 128 B], swizzled like a TMA tile
  float* recv = reinterpret_cast<float*>(smem + p.ring_bytes + p.xs_bytes);  // rank 0: [SK - 1][parts][NC][128 rows] partial sums
  __shared__ uint64_t full[SK_MAX_STAGES], empty[SK_MAX_STAGES], acc_full[2], acc_empty[2], x_ready, part_full;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rank = p.SK > 1 ? static_cast<int>(cluster_ctarank()) : 0;
  const int cluster_id = static_cast<int>(blockIdx.x) / p.SK, n_clusters = static_cast<int>(gridDim.x) / p.SK;

  if (tid == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&acc_full[a], 1);
      mbar_init(&acc_empty[a], 128);
    }
    mbar_init(&x_ready, SK_NORM_WARPS * 32);
    mbar_init(&part_full, p.SK > 1 ? (p.SK - 1) * 128 : 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmX);
  }
  if (warp == 1) {
    tmem_alloc(&tmem_base_s, p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (p.SK > 1) cluster_sync_all();  // rank 0's part_full is initialised before any peer can arrive on it
  pdl_trigger();                     // the next kernel of the step may take this SM's resources as soon as this CTA retires
  const uint32_t tmem = tmem_base_s;
  const int kb_lo = static_cast<int>(static_cast<long long>(p.kblocks) * rank / p.SK);
  const int kb_hi = static_cast<int>(static_cast<long long>(p.kblocks) * (rank + 1) / p.SK);
  const int n_ks = (kb_hi - kb_lo + p.KS - 1) / p.KS;  // ring stages per tile
  const int my_tiles = (p.n_tiles - cluster_id + n_clusters - 1) / n_clusters;
  const int total = my_tiles * n_ks;
  const uint32_t acc_cols = static_cast<uint32_t>(p.parts * p.NC);  // TMEM columns of one accumulator buffer

  if (warp == 0) {
    // ======================= TMA producer =======================
    if (lane == 0) {
      // ring position / tile position are carried incrementally: a 64-bit division per stage on one thread costs more
      // than the stage's loads
      struct Pos {
        int s, phase, ti, ks;
      };
      auto advance = [&](Pos& q) {
        if (++q.s == p.stages) q.s = 0, q.phase ^= 1;
        if (++q.ks == n_ks) q.ks = 0, ++q.ti;
      };
      auto issue = [&](const Pos& q, bool weights, bool inputs) {
        const int tile = cluster_id + q.ti * n_clusters;
        const int kb0 = kb_lo + q.ks * p.KS;
        uint8_t* st = ring + static_cast<uint32_t>(q.s) * p.stage_bytes;
        if (weights) {
          mbar_wait(&empty[q.s], q.phase ^ 1);
          // one box per operand and stage: KS consecutive k-blocks land as KS canonical [rows x 128 B] tiles (k-blocks past
          // K are zero-filled and still counted)
          mbar_expect_tx(&full[q.s], p.stage_bytes);
          tma_load_4d(st, &tmW, &full[q.s], 0, tile * p.R, 0, kb0);
        }
        if (inputs && !p.norm)
          tma_load_3d(st + static_cast<uint32_t>(p.KS * p.parts) * p.w_tile_bytes, &tmX, &full[q.s], 0, 0, kb0);
      };
      // the weights do not depend on the previous kernel: fill the ring, then wait for it, then fetch the inputs
      const int pre = total < p.stages ? total : p.stages;
      Pos q = {0, 0, 0, 0};
      for (int i = 0; i < pre; ++i, advance(q)) issue(q, true, false);
      pdl_wait();
      Pos q2 = {0, 0, 0, 0};
      for (int i = 0; i < pre; ++i, advance(q2)) issue(q2, false, true);
      for (int it = pre; it < total; ++it, advance(q)) issue(q, true, true);
    }
  } else if (warp == 1) {
    // ======================= MMA issuer =======================
    const uint32_t idesc = umma_idesc_bf16_f32(128, p.NC, 0, 0);
    if (p.norm) {
      mbar_wait(&x_ready, 0);
      tc_fence_after();
    }
    int s = 0;
    uint32_t phase = 0;
    for (int ti = 0; ti < my_tiles; ++ti) {
      const uint32_t a = ti & 1;
      mbar_wait(&acc_empty[a], ((ti >> 1) & 1) ^ 1);
      tc_fence_after();
      for (int ks = 0; ks < n_ks; ++ks) {
        const int kb0 = kb_lo + ks * p.KS, nkb = min(p.KS, kb_hi - kb0);
        mbar_wait(&full[s], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t st = smem_u32(ring) + static_cast<uint32_t>(s) * p.stage_bytes;
          for (int j = 0; j < nkb && !p.debug_no_mma; ++j) {
            const uint32_t xa = p.norm ? smem_u32(xs) + static_cast<uint32_t>(kb0 - kb_lo + j) * p.x_tile_bytes
                                       : st + static_cast<uint32_t>(p.KS * p.parts) * p.w_tile_bytes + static_cast<uint32_t>(j) * p.x_tile_bytes;
            for (int q = 0; q < p.parts; ++q) {
              const uint32_t wa = st + static_cast<uint32_t>(j * p.parts + q) * p.w_tile_bytes;
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                umma_ss(tmem + a * acc_cols + q * p.NC, umma_smem_desc_sw128(wa + kk * 32, 16, 1024),
                        umma_smem_desc_sw128(xa + kk * 32, 16, 1024), idesc, (ks > 0 || j > 0 || kk > 0) ? 1u : 0u);
            }
          }
          umma_commit(&empty[s]);
          if (ks == n_ks - 1) umma_commit(&acc_full[a]);
        }
        __syncwarp();
        if (++s == p.stages) s = 0, phase ^= 1;
      }
    }
  } else {
    // ======================= warps 2-7: RMSNorm prologue; warps 4-7: epilogue =======================
    if (p.norm) {
      pdl_wait();
      // a warp per input row, the row's 16-byte chunks in registers (K <= 2048: one round trip to L2; beyond, the second pass re-reads L1 / L2); the statistics
      // cover the whole row, the staged copy this CTA's K slice
      const int w2 = warp - 2;
      const int nvec = p.K / 8;
      constexpr int CH = 8;
      const bool one_seg = nvec <= 32 * CH;
      const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
      for (int b = w2; b < p.B; b += SK_NORM_WARPS) {
        const uint4* row = reinterpret_cast<const uint4*>(p.x + static_cast<long long>(b) * p.ldx);
        uint4 a[CH];
        float ss = 0.f;
        for (int v0 = 0; v0 < nvec; v0 += 32 * CH) {
#pragma unroll
          for (int u = 0; u < CH; ++u) {
            const int v = v0 + u * 32 + lane;
            a[u] = v < nvec ? row[v] : zero4;
          }
#pragma unroll
          for (int u = 0; u < CH; ++u) {
            float f[8];
            unpack8(a[u], f);
#pragma unroll
            for (int e = 0; e < 8; ++e) ss = fmaf(f[e], f[e], ss);
          }
        }
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        const float rstd = rsqrtf(ss / static_cast<float>(p.K) + p.eps);
        for (int v0 = 0; v0 < nvec; v0 += 32 * CH) {
#pragma unroll
          for (int u = 0; u < CH; ++u) {
            const int v = v0 + u * 32 + lane;
            const int kb = v >> 3;
            if (v >= nvec || kb < kb_lo || kb >= kb_hi) continue;
            float f[8], nw[8];
            unpack8(one_seg ? a[u] : row[v], f);
            unpack8(reinterpret_cast<const uint4*>(p.norm_w)[v], nw);
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = nw[e] * bf_round(f[e] * rstd);  // LlamaRMSNorm.forward
            *reinterpret_cast<uint4*>(xs + static_cast<size_t>(kb - kb_lo) * p.x_tile_bytes + sw128_offset(b, v & 7)) = pack8(f);
          }
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(&x_ready);
    }
    if (warp >= 4) {
      pdl_wait();
      const int quarter = warp & 3;
      const int row = quarter * 32 + lane;  // TMEM lane = weight row of the tile
      const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
      uint32_t r[16], r2[16];
      for (int ti = 0; ti < my_tiles; ++ti) {
        const uint32_t a = ti & 1;
        const int tile = cluster_id + ti * n_clusters;
        const int n = tile * p.R + row;  // output column (pair: index of the gate / up pair)
        const bool ok = row < p.R && n < p.units;
        mbar_wait(&acc_full[a], (ti >> 1) & 1);
        tc_fence_after();
        for (int c = 0; c < p.NC / 16; ++c) {  // 16 prompts per pass
          tmem_ld_32x32b_x16(tmem + lane_addr + a * acc_cols + c * 16, r);
          if (p.parts == 2) tmem_ld_32x32b_x16(tmem + lane_addr + a * acc_cols + p.NC + c * 16, r2);
          tmem_ld_wait();
          if (c == p.NC / 16 - 1) {  // the accumulator is in registers: the next tile's MMAs may overwrite it
            tc_fence_before();
            mbar_arrive(&acc_empty[a]);
          }
          const int nb = min(16, p.B - c * 16);
          if (p.SK > 1) {
            // K-split: partial sums of ranks 1.. travel to rank 0 ([rank - 1][part][prompt][row], coalesced over rows)
            if (rank > 0) {
              const uint32_t base = mapa_shared(smem_u32(recv), 0) +
                                    static_cast<uint32_t>((((rank - 1) * p.parts) * p.NC + c * 16) * 128 + row) * 4u;
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                if (i < nb) {
                  st_cluster_f32(base + static_cast<uint32_t>(i * 128) * 4u, __uint_as_float(r[i]));
                  if (p.parts == 2) st_cluster_f32(base + static_cast<uint32_t>((p.NC + i) * 128) * 4u, __uint_as_float(r2[i]));
                }
              }
              if (c == p.NC / 16 - 1) mbar_arrive_release_cluster(mapa_shared(smem_u32(&part_full), 0));
              continue;
            }
            if (c == 0) mbar_wait_acquire_cluster(&part_full, ti & 1);
            for (int rk = 1; rk < p.SK; ++rk) {  // fixed order: the sum does not depend on arrival order
              const float* src = recv + (((rk - 1) * p.parts) * p.NC + c * 16) * 128 + row;
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                if (i < nb) {
                  r[i] = __float_as_uint(__uint_as_float(r[i]) + src[i * 128]);
                  if (p.parts == 2) r2[i] = __float_as_uint(__uint_as_float(r2[i]) + src[(p.NC + i) * 128]);
                }
              }
            }
          }
          if (!ok) continue;
          const float bias = p.bias ? __bfloat162float(p.bias[n]) : 0.f;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            if (i >= nb) continue;
            const int b = c * 16 + i;
            float v;
            if (p.parts == 2) {
              const float g = bf_round(__uint_as_float(r[i])), u = bf_round(__uint_as_float(r2[i]));
              v = bf_round(g / (1.f + __expf(-g))) * u;  // LlamaMLP: bf16(bf16(silu(gate)) * up)
            } else {
              v = bf_round(__uint_as_float(r[i]) + bias);  // nn.Linear output in bf16
              if (p.resid) v = bf_round(v + __bfloat162float(p.resid[static_cast<long long>(b) * p.ldr + n]));  // residual + hidden
            }
            p.out[static_cast<long long>(b) * p.ldo + n] = __float2bfloat16(v);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, p.tmem_cols);
}

int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return (e && *e) ? atoi(e) : dflt;
}
// bring-up knobs, read once per process
struct SkinnyKnobs {
  int no_mma, smem_kb, debug, sk, r, ks, stages;
};
const SkinnyKnobs& knobs() {
  static const SkinnyKnobs k = {env_int("SDAG_SKINNY_NOMMA", 0), env_int("SDAG_SKINNY_SMEM_KB", 0), env_int("SDAG_SKINNY_DEBUG", 0),
                                env_int("SDAG_SKINNY_SK", 0),    env_int("SDAG_SKINNY_R", 0),       env_int("SDAG_SKINNY_KS", 0),
                                env_int("SDAG_SKINNY_STAGES", 0)};
  return k;
}

// launch_pdl with a thread-block cluster of `cluster` CTAs along x
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster, Args... args) {
  static int use_pdl = -1;
  if (use_pdl < 0) {
    const char* e = getenv("SDAG_NO_PDL");
    use_pdl = (e && atoi(e)) ? 0 : 1;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (cluster > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = static_cast<unsigned>(cluster);
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (use_pdl) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace

extern "C" int llm_skinny_linear(const void* x, int64_t ldx, const void* weight, const void* bias, const void* norm_weight, float eps,
                                 const void* residual, int64_t ldr, void* out, int64_t ldo, int B, int N, int K, int pair_swiglu,
                                 void* stream) {
  if (!x || !weight || !out || B <= 0 || N <= 0 || K <= 0) return SDAG_ERR_INVALID;
  if (B > 32 || (K % 64) || (N % 8) || (ldx % 8) || (ldo % 8) || (residual && (ldr % 8))) return SDAG_ERR_UNSUPPORTED;
  if (pair_swiglu && (bias || residual || (N % 16))) return SDAG_ERR_INVALID;
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(weight) | reinterpret_cast<uintptr_t>(out) |
       reinterpret_cast<uintptr_t>(bias) | reinterpret_cast<uintptr_t>(residual) | reinterpret_cast<uintptr_t>(norm_weight)) & 15)
    return SDAG_ERR_INVALID;
  const int n_sm = device_sm_count();
  SkinnyParams p = {};
  p.x = static_cast<const __nv_bfloat16*>(x);
  p.ldx = ldx;
  p.norm_w = static_cast<const __nv_bfloat16*>(norm_weight);
  p.eps = eps;
  p.bias = static_cast<const __nv_bfloat16*>(bias);
  p.resid = static_cast<const __nv_bfloat16*>(residual);
  p.ldr = ldr;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.ldo = ldo;
  p.B = B, p.K = K, p.N = N;
  p.RB = (B + 7) / 8 * 8;
  p.NC = B <= 16 ? 16 : 32;
  p.parts = pair_swiglu ? 2 : 1;
  p.half = N / 2;
  p.units = pair_swiglu ? p.half : N;
  p.norm = norm_weight ? 1 : 0;
  p.kblocks = K / 64;
  p.x_tile_bytes = static_cast<uint32_t>(p.RB) * 128u;
  p.debug_no_mma = knobs().no_mma;

  // Decomposition: SK CTAs (a cluster) share a row tile and split K; rows per tile R <= 128.  Model per choice: the stream
  // (bytes per CTA at the smaller of the HBM share and ~64 GB/s per SM) against the MMA time (~56 cycles per 128-row x 16-k
  // slice, whatever R is), plus a fixed cost for the cluster reduction.
  // Shared memory per CTA: layers up to ~150 MB of weights take at most half an SM (two CTAs per SM — of this grid or of two
  // consecutive kernels of the step, so that the next layer's ring fills while this one drains); the large ones (gate/up,
  // LM head) a whole SM and a deeper ring.
  constexpr uint32_t SMEM_MAX = 224 * 1024;
  const int forced_kb = knobs().smem_kb;
  int cls = forced_kb > 0 ? (forced_kb <= 112 ? 0 : 1) : (static_cast<double>(N) * K * 2.0 < 150e6 ? 0 : 1);
  uint32_t SMEM_BUDGET = forced_kb > 0 ? static_cast<uint32_t>(forced_kb) * 1024u : (cls == 0 ? 110u * 1024u : SMEM_MAX);
  if (SMEM_BUDGET > SMEM_MAX) return SDAG_ERR_INVALID;
  static bool configured[SDAG_MAX_DEVICES] = {};
  static int max_clusters[SDAG_MAX_DEVICES][2][SK_MAX_CLUSTER + 1] = {};  // co-resident clusters per size (0: not asked yet, -1: none)
  const int dev = device_slot();
  if (dev < 0) return SDAG_ERR_UNSUPPORTED;
  if (!configured[dev]) {
    SDAG_CUDA_TRY(cudaFuncSetAttribute(skinny_linear_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(SMEM_MAX)));
    configured[dev] = true;
  }
  int ctas_per_sm = SMEM_BUDGET <= 112u * 1024u ? 2 : 1;
  auto clusters_for = [&](int sk) -> int {  // a cluster is placed inside one GPC: fewer than slots / sk fit at once
    if (sk == 1) return n_sm * ctas_per_sm;
    if (max_clusters[dev][cls][sk] == 0) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(static_cast<unsigned>(n_sm * ctas_per_sm / sk * sk));
      cfg.blockDim = dim3(SK_THREADS);
      cfg.dynamicSmemBytes = SMEM_BUDGET;
      cudaLaunchAttribute at;
      at.id = cudaLaunchAttributeClusterDimension;
      at.val.clusterDim.x = static_cast<unsigned>(sk);
      at.val.clusterDim.y = 1;
      at.val.clusterDim.z = 1;
      cfg.attrs = &at;
      cfg.numAttrs = 1;
      int n = 0;
      const cudaError_t e = cudaOccupancyMaxActiveClusters(&n, skinny_linear_kernel, &cfg);
      if (e != cudaSuccess) {
        (void)cudaGetLastError();
        n = 0;
      }
      if (knobs().debug)
        fprintf(stderr, "llm_skinny: cluster size %d, %u KB -> %d co-resident clusters (%s)\n", sk, SMEM_BUDGET / 1024u, n, cudaGetErrorString(e));
      max_clusters[dev][cls][sk] = n > 0 ? n : -1;
    }
    return max_clusters[dev][cls][sk];
  };
  int best_sk = 1, best_r = 8;
  double best_t = 1e30;
  const int forced_sk = knobs().sk;
  for (int attempt = 0; attempt < 2 && best_t > 1e29; ++attempt) {
  if (attempt == 1) {  // nothing fits half an SM (a long K staged for the RMSNorm prologue): take a whole one
    if (cls == 1 || forced_kb > 0) break;
    cls = 1, SMEM_BUDGET = SMEM_MAX, ctas_per_sm = 1;
  }
  for (int sk = 1; sk <= SK_MAX_CLUSTER; ++sk) {
    if (forced_sk > 0 && sk != forced_sk) continue;
    if (sk > 1 && (B > 16 || p.kblocks / sk < 4)) continue;
    const int groups = clusters_for(sk);
    if (groups < 1) continue;
    const int rounds = (p.units + groups * 128 - 1) / (groups * 128);
    if (sk > 1 && rounds > 1) continue;
    int r = (p.units + groups * rounds - 1) / (groups * rounds);
    r = (r + 7) / 8 * 8;
    if (r > 128) r = 128;
    {  // does it fit?  staged input + partial sums + over-read pad + two one-k-block stages
      const uint32_t xs_b = p.norm ? static_cast<uint32_t>((p.kblocks + sk - 1) / sk) * p.x_tile_bytes : 0u;
      const uint32_t rv_b = sk > 1 ? static_cast<uint32_t>((sk - 1) * p.parts * p.NC * 128 * 4) : 0u;
      const uint32_t w_b = static_cast<uint32_t>(r) * 128u;
      const uint32_t ov = static_cast<uint32_t>(SK_A_SPAN) - w_b;
      const uint32_t pd = ov > xs_b + rv_b ? ov - xs_b - rv_b + 1024u : 1024u;
      if (SMEM_BUDGET < 1024u + xs_b + rv_b + pd + 2u * (p.parts * w_b + (p.norm ? 0u : p.x_tile_bytes))) continue;
    }
    const int tiles = (p.units + r - 1) / r;
    const int active = (tiles < groups ? tiles : groups) * sk;
    const double bytes_cta = static_cast<double>(rounds) * r * p.parts * K * 2.0 / sk;
    const double share = active > n_sm ? 2.0 : 1.0;  // two CTAs on an SM share its tensor pipe and its load bandwidth
    const double per_cta_bw = 6.4e12 / active < 64e9 / share ? 6.4e12 / active : 64e9 / share;
    const double t_stream = bytes_cta / per_cta_bw;
    const double t_mma = static_cast<double>(rounds) * (p.kblocks / static_cast<double>(sk)) * p.parts * 4.0 * 56.0 * share / 1.2e9;
    const double t = (t_stream > t_mma ? t_stream : t_mma) + (sk > 1 ? 1.0e-6 : 0.0);
    if (t < best_t - 1e-9) best_t = t, best_sk = sk, best_r = r;
  }
  }
  if (best_t > 1e29) return SDAG_ERR_UNSUPPORTED;  // e.g. too many input rows x K to stage: normalise first (llm_rmsnorm)
  p.SK = best_sk;
  p.R = best_r;
  const int forced_r = knobs().r;
  if (forced_r >= 8 && forced_r <= 128 && forced_r % 8 == 0) p.R = forced_r;
  p.n_tiles = (p.units + p.R - 1) / p.R;
  p.w_tile_bytes = static_cast<uint32_t>(p.R) * 128u;
  const int slice_max = (p.kblocks + p.SK - 1) / p.SK;
  p.xs_bytes = p.norm ? static_cast<uint32_t>(slice_max) * p.x_tile_bytes : 0u;
  p.recv_bytes = p.SK > 1 ? static_cast<uint32_t>((p.SK - 1) * p.parts * p.NC * 128 * 4) : 0u;
  // k-blocks per stage: ~24 KB of weights per box, preferring a count that wastes no k-blocks at the end of the K slice
  const uint32_t per_kb = static_cast<uint32_t>(p.parts) * p.w_tile_bytes + (p.norm ? 0u : p.x_tile_bytes);
  int ks_target = static_cast<int>((24u * 1024u + per_kb / 2) / per_kb);
  if (ks_target < 1) ks_target = 1;
  if (ks_target > 8) ks_target = 8;
  const int slice_min = p.kblocks / p.SK;
  if (ks_target > slice_min) ks_target = slice_min;
  // behind the ring: xs, recv, and whatever is still missing for the A over-read of the last weight tile (the x tiles' own
  // N = 16 over-read needs 1 KB)
  const uint32_t over = static_cast<uint32_t>(SK_A_SPAN) - p.w_tile_bytes;
  const uint32_t pad = over > p.xs_bytes + p.recv_bytes ? over - p.xs_bytes - p.recv_bytes + 1024u : 1024u;
  if (SMEM_BUDGET < 1024u + p.xs_bytes + p.recv_bytes + pad + 2u * per_kb) return SDAG_ERR_UNSUPPORTED;
  const uint32_t ring_budget = SMEM_BUDGET - 1024u - p.xs_bytes - p.recv_bytes - pad;
  while (ks_target > 1 && 3u * ks_target * per_kb > ring_budget) --ks_target;  // at least three stages in the ring
  p.KS = ks_target;
  {
    int best_waste = 1 << 30;
    for (int ks = ks_target; ks >= (ks_target + 1) / 2; --ks) {
      const int waste = (slice_max + ks - 1) / ks * ks - slice_max;
      if (waste < best_waste) best_waste = waste, p.KS = ks;
    }
  }
  const int forced_ks = knobs().ks;
  if (forced_ks >= 1 && forced_ks <= 8 && forced_ks <= slice_min) p.KS = forced_ks;
  p.stage_bytes = static_cast<uint32_t>(p.KS) * per_kb;
  p.stages = static_cast<int>(ring_budget / p.stage_bytes);
  const int forced_s = knobs().stages;
  if (forced_s > 0 && forced_s < p.stages) p.stages = forced_s;
  if (p.stages > SK_MAX_STAGES) p.stages = SK_MAX_STAGES;
  if (p.stages < 2) return SDAG_ERR_UNSUPPORTED;
  uint32_t cols = 32;
  while (cols < 2u * static_cast<uint32_t>(p.parts * p.NC)) cols <<= 1;
  p.tmem_cols = cols;
  p.ring_bytes = static_cast<uint32_t>(p.stages) * p.stage_bytes;
  const size_t smem = 1024 + static_cast<size_t>(p.ring_bytes) + p.xs_bytes + p.recv_bytes + pad;

  // weights as {64 k, rows of a part, parts (gate | up), k-blocks}: a box of KS k-blocks lands as KS x parts [R x 128 B]
  // tiles; inputs as {64 k, prompts, k-blocks}
  CUtensorMap tmW, tmX;
  {
    const uint64_t rows = static_cast<uint64_t>(p.units);
    const uint64_t dims[4] = {64, rows, static_cast<uint64_t>(p.parts), static_cast<uint64_t>(p.kblocks)};
    const uint64_t strides[3] = {static_cast<uint64_t>(K) * 2, rows * static_cast<uint64_t>(K) * 2, 128};
    const uint32_t box[4] = {64, static_cast<uint32_t>(p.R), static_cast<uint32_t>(p.parts), static_cast<uint32_t>(p.KS)};
    if (make_tmap(&tmW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, weight, dims, strides, box)) return SDAG_ERR_INVALID;
  }
  {
    const uint64_t dims[3] = {64, static_cast<uint64_t>(B), static_cast<uint64_t>(p.kblocks)};
    const uint64_t strides[2] = {static_cast<uint64_t>(ldx) * 2, 128};
    const uint32_t box[3] = {64, static_cast<uint32_t>(p.RB), static_cast<uint32_t>(p.KS)};
    if (make_tmap(&tmX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, x, dims, strides, box)) return SDAG_ERR_INVALID;
  }
  const int groups = clusters_for(p.SK);
  if (knobs().debug)
    fprintf(stderr, "llm_skinny: B=%d N=%d K=%d pair=%d norm=%d -> SK=%d R=%d tiles=%d KS=%d stages=%d stage=%u B smem=%zu groups=%d\n", B, N, K,
            p.parts == 2, p.norm, p.SK, p.R, p.n_tiles, p.KS, p.stages, p.stage_bytes, smem, groups);
  if (smem > SMEM_BUDGET || groups < 1 || (p.SK > 1 && p.n_tiles > groups)) return SDAG_ERR_UNSUPPORTED;
  const int grid = (p.n_tiles < groups ? p.n_tiles : groups) * p.SK;
  return static_cast<int>(launch_pdl_cluster(skinny_linear_kernel, dim3(grid), dim3(SK_THREADS), smem, static_cast<cudaStream_t>(stream),
                                             p.SK, tmW, tmX, p));
}
