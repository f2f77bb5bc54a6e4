// DeepSDF decoder on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a only.
//
// replaces: the nine nn.Linear GEMMs of Decoder.forward
//           (sdfrenderer/deepsdf/networks/deep_sdf_decoder_scale.py:88-94) and the
//           autograd pass that yields the normals (sdfrenderer/grid.py:55-56), as one
//           persistent kernel: forward, then the input-gradient backward, per 64-point tile.
//
// Orientation.  A 128-point x 512-wide fp32 activation tile does not fit on chip next to
// its successor (SURVEY.md "SMEM/TMEM budget"), so the GEMM is issued transposed:
//     D^T[feature, point] = W[feature, k] * H^T[k, point]
// A = weight tile (128 features x 32 k, K-major, streamed from L2 with cp.async.bulk into a
// 5-stage ring), B = the CTA's activations (64 points x 512 k, resident in shared memory,
// MN-major so that the 8 points one epilogue thread packs are one 16-byte store),
// D = 4 x (128 lanes x 128 columns) fp32 accumulators in TMEM.  TMEM lane == output feature,
// so the epilogue thread that owns a lane applies bias / ReLU / mask for one feature across
// the 64 points and writes the next layer's B operand in place.
//
// Precision.  fp32 parity (1e-4 on normals) needs more than one TF32/BF16 pass (SURVEY.md
// "Tensor-core precision").  Every operand is split into two fp16 halves x = hi + lo
// (22 significand bits, power-of-two pre-scaling keeps lo out of the subnormals) and each
// product is  hi*hi + hi*lo + lo*hi  with fp32 accumulation in TMEM.  Per K=16 step that is
// two MMAs: W_hi x [H_hi ; H_lo] (N = 128: columns 0-63 collect hi*hi, 64-127 hi*lo) and
// W_lo x H_hi (N = 64) accumulated onto columns 64-127.  Keeping the 2^-11-sized cross terms
// in their own accumulator shortens the (truncating) accumulation chain of the main sum from
// 96 to 32 adds; the epilogue adds the two columns in fp32.  Half the operand bytes of 3xTF32
// and twice its MMA rate.
//
// Warp roles (320 threads): warps 0-7 epilogue (TMEM lane quadrant = warp & 3, point half =
// warp >> 2), warp 8 weight producer (one lane), warp 9 MMA issuer (one elected lane) + TMEM allocator.
#include <cuda_fp16.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "trace.cuh"

namespace sdfr {

namespace {

constexpr int NPTS = 64;                 // points per CTA tile
constexpr int KC = 32;                   // k per weight stage (two K=16 MMA steps)
constexpr int TILE_BYTES = 16384;        // 128 x 32 fp16 hi + the same lo
constexpr int TILE_HALF_BYTES = 8192;
// weight-ring depth of mlp_tc_kernel<NP>: the smaller the point tile, the more shared memory is left for stages
template <int NP> struct RingDepth { static constexpr int value = NP <= 16 ? 11 : 5; };
// B operand (activations), MN-major, no swizzle: core matrix = 8 k-rows x (8 points = 16 B).
// Per 8-k chunk: 8 point groups of the hi halves (1024 B) followed by 8 point groups of the lo
// halves (1024 B), so one descriptor with N = 128 covers [hi ; lo] and N = 64 covers hi only.
constexpr int B_CHUNK = 2048;
constexpr int A_LBO = 2048;              // A (K-major): core matrices adjacent in K are 16 row groups apart
constexpr int A_SBO = 128;
constexpr int B_SBO = 128;               //               next 8-point group
constexpr int TMEM_COLS = 512;           // 4 M-blocks x (64 main + 64 cross-term) columns
constexpr int NTHREADS = 320;          // 8 epilogue warps + producer + MMA issuer
constexpr int NEPI = 256;
constexpr int MAX_TC_LAYERS = 9;
constexpr float ACT_SCALE = 32.f;        // forward activations are stored as h * 2^5
constexpr float BWD_SCALE = 256.f;       // backward deltas as delta * 2^8

struct TcPassDev {
  int kind;          // 0 forward hidden, 1 forward last, 2 backward hidden, 3 backward first
  int layer;
  int m_blocks, k_chunks;
  int rows;          // logical output rows of this pass
  int cat_dim, cat_off;   // rows [rows, rows+cat_dim) are the concatenated input (forward) / its gradient (backward)
  int prev_rows;     // backward: fan-out of layer-1 (rows below it go through that layer's ReLU mask)
  float inv_scale;   // 1 / (weight scale * input activation scale)
  float out_scale;   // scale applied to what the epilogue writes back as the next B operand
  long long tile0;
  const float* bias; // forward: [rows]
};

struct TcTable {
  int num_passes, num_layers, in0, latent, use_tanh;
  int last_k;                    // fan-in of the last Linear
  float last_wscale_inv;
  const float* last_w;           // [last_k] un-scaled last-layer weights (backward of the last Linear is an outer product)
  long long tiles_per_point_tile;
  TcPassDev pass[2 * MAX_TC_LAYERS];
};

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], kind::f16, fp32 accumulate
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the mbarrier once every MMA issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// K-major, no-swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, version 1):
// 8-row x 16-byte core matrices; LBO = byte distance between core matrices adjacent in K,
// SBO = between core matrices adjacent in M/N.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
  return d;                 // base offset 0, LBO mode 0, layout type 0 = SWIZZLE_NONE
}

// instruction descriptor (cute::UMMA::InstrDescriptor): D fp32, A/B fp16, A K-major, B MN-major, M=128
__host__ __device__ constexpr uint32_t idesc_for(int n) {
  return (1u << 4) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
constexpr uint32_t kIdesc64 = idesc_for(64);

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// coarse pass: only the hi halves are ever read by the MMAs
__device__ __forceinline__ void pack8_store_hi(unsigned char* chunk_row, int pg, const float (&h)[8]) {
  uint4 hi;
  uint32_t* hw = reinterpret_cast<uint32_t*>(&hi);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __half2 a = __floats2half2_rn(h[2 * i], h[2 * i + 1]);
    hw[i] = *reinterpret_cast<const uint32_t*>(&a);
  }
  *reinterpret_cast<uint4*>(chunk_row + pg * 128) = hi;
}

struct SmemPlan {
  uint32_t stages, b, inp, dinp, g, bars, tmem_slot, total;
};
// NP = points per CTA tile: 64 for throughput, 16 to spread a short row list (the band pass of a
// single detection: ~1 850 rows) over enough CTAs to fill the machine.
template <int NP>
__host__ __device__ inline SmemPlan make_plan(int num_layers, int in0) {
  SmemPlan p;
  uint32_t o = 0;
  p.stages = o; o += RingDepth<NP>::value * TILE_BYTES;
  p.b = o; o += 64 * (NP * 32);            // 64 k-chunks x (hi + lo) point groups
  const uint32_t in_pad = (uint32_t)((in0 + 7) & ~7);
  p.inp = o; o += in_pad * NP * 4;
  p.dinp = o; o += in_pad * NP * 4;
  p.g = o; o += 64 * 4;
  p.bars = o; o += 32 * 8;
  p.tmem_slot = o; o += 16;
  p.total = o;
  return p;
}

#ifdef SDFR_TC_PROFILE
__device__ unsigned long long g_tc_prof[16];
#define PROF_T0() const long long _t0 = clock64()
#define PROF_ADD(slot) do { if (blockIdx.x == 0) atomicAdd(&g_tc_prof[slot], (unsigned long long)(clock64() - _t0)); } while (0)
#else
#define PROF_T0() do {} while (0)
#define PROF_ADD(slot) do {} while (0)
#endif

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}

__device__ __forceinline__ void tmem_ldg(uint32_t taddr, uint32_t (&v)[8]) { tmem_ld8(taddr, v); }
__device__ __forceinline__ void tmem_ldg(uint32_t taddr, uint32_t (&v)[16]) { tmem_ld16(taddr, v); }

// hi halves at row + pg*128, lo halves LO bytes further
template <int LO>
__device__ __forceinline__ float pack8_store_t(unsigned char* chunk_row, int pg, const float (&h)[8]) {
  uint4 hi, lo;
  uint32_t* hw = reinterpret_cast<uint32_t*>(&hi);
  uint32_t* lw = reinterpret_cast<uint32_t*>(&lo);
  float amax = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __half2 a = __floats2half2_rn(h[2 * i], h[2 * i + 1]);
    const float2 b = __half22float2(a);
    const __half2 l = __floats2half2_rn(h[2 * i] - b.x, h[2 * i + 1] - b.y);
    hw[i] = *reinterpret_cast<const uint32_t*>(&a);
    lw[i] = *reinterpret_cast<const uint32_t*>(&l);
    amax = fmaxf(amax, fmaxf(fabsf(h[2 * i]), fabsf(h[2 * i + 1])));
  }
  *reinterpret_cast<uint4*>(chunk_row + pg * 128) = hi;
  *reinterpret_cast<uint4*>(chunk_row + LO + pg * 128) = lo;
  return amax;
}

template <int NP>
__global__ void __launch_bounds__(NTHREADS, 1)
mlp_tc_kernel(const TcTable* __restrict__ tabp, const unsigned char* __restrict__ tiles, MlpInputs in,
              float* __restrict__ sdf_out, float* __restrict__ dinput_out, int* __restrict__ overflow_flag,
              unsigned long long* __restrict__ mask_scratch, long long num_point_tiles, int group) {
  extern __shared__ __align__(1024) unsigned char smem[];
  if (in.count_dev && *in.count_dev <= 0) return;   // nothing to evaluate (uniform over the whole grid)
  const TcTable& T = *tabp;
  const int num_layers = T.num_layers, in0 = T.in0, latent = T.latent, use_tanh = T.use_tanh;
  constexpr int NSTAGE = RingDepth<NP>::value;
  constexpr int BCH = NP * 32;            // bytes per 8-k chunk of B: NP/8 hi groups then NP/8 lo groups
  constexpr int LO = NP * 16;             // offset of the lo groups inside a chunk
  constexpr int PT = NP / 2;              // points per epilogue thread (the two point halves)
  constexpr int GW = PT >= 16 ? 16 : 8;   // points per TMEM load group
  constexpr int G = PT / GW;              // load groups per thread
  constexpr int PG = GW / 8;              // 8-point packs (16-byte stores) per load group
  constexpr int TCOLS = 8 * NP < 32 ? 32 : 8 * NP;   // 4 M-blocks x (NP main + NP cross) columns
  constexpr uint32_t kIdMain = idesc_for(2 * NP), kIdCross = idesc_for(NP);
  const SmemPlan P = make_plan<NP>(num_layers, in0);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  unsigned char* bop = smem + P.b;
  // ReLU sign words (one u64 per feature per hidden layer) live in an L2-resident global scratch:
  // written and read back by the same thread, 32 KB per CTA, which buys two more weight stages.
  unsigned long long* masks = mask_scratch + (size_t)blockIdx.x * (size_t)(num_layers - 1) * 512;
  float* inp = reinterpret_cast<float*>(smem + P.inp);     // [in_pad][64]
  float* dinp = reinterpret_cast<float*>(smem + P.dinp);
  float* gbuf = reinterpret_cast<float*>(smem + P.g);
  const uint32_t bars = smem_u32(smem + P.bars);
  const uint32_t bar_full = bars, bar_empty = bars + 8 * NSTAGE, bar_acc = bars + 8 * (2 * NSTAGE),
                 bar_act = bars + 8 * (2 * NSTAGE + 1);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + P.tmem_slot);
  const int in_pad = (in0 + 7) & ~7;
  const bool want_grad = dinput_out != nullptr;
  const int npass = want_grad ? T.num_passes : num_layers;   // forward passes come first
  const long long n_rows = mlp_rows(in);           // the row count may live on the device (band pass)
  if (in.count_dev) num_point_tiles = (n_rows + NP - 1) / NP;

  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    mbar_init(bar_acc, 1);
    mbar_init(bar_act, NEPI);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 9) tmem_alloc(smem_u32(smem + P.tmem_slot), TCOLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // point tiles blockIdx.x, blockIdx.x + gridDim.x, ...
  long long my_tiles = 0;
  for (long long pt = blockIdx.x; pt < num_point_tiles; pt += gridDim.x) ++my_tiles;

  if (warp == 8) {
    // ===================== weight producer =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (long long it = 0; it < my_tiles; ++it) {
        for (int p = 0; p < npass; ++p) {
          const long long n = (long long)T.pass[p].m_blocks * T.pass[p].k_chunks;
          const unsigned char* src = tiles + T.pass[p].tile0 * TILE_BYTES;
          for (long long t = 0; t < n; ++t) {
            { PROF_T0(); mbar_wait(bar_empty + 8 * stage, phase ^ 1); PROF_ADD(0); }
            mbar_expect_tx(bar_full + 8 * stage, TILE_BYTES);
            bulk_g2s(smem_u32(smem + P.stages + stage * TILE_BYTES), src + t * TILE_BYTES, TILE_BYTES, bar_full + 8 * stage);
            if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 9) {
    // ===================== MMA issuer =====================
    // All 32 lanes run the loop with identical (warp-uniform) values so that the descriptor
    // arithmetic stays on the uniform datapath; one elected lane issues the tcgen05 instructions.
    // (A single thread executing ~80 dependent scalar instructions per K step takes ~240 cycles,
    // 2.5x the 96 cycles of tensor work it feeds - measured with tools/umma_bench.cu.)
    uint32_t stage = 0, phase = 0, act_phase = 0;
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint64_t desc_a_base =
        make_desc(smem_u32(smem + P.stages), A_LBO, A_SBO);             // + (stage*TILE_BYTES + ...) >> 4
    const uint64_t desc_b_base = make_desc(smem_u32(bop), BCH, B_SBO);
    for (long long it = 0; it < my_tiles; ++it) {
      for (int p = 0; p < npass; ++p) {
        const int m_blocks = __ldg(&T.pass[p].m_blocks), k_chunks = __ldg(&T.pass[p].k_chunks);
        { PROF_T0(); mbar_wait(bar_act, act_phase); if (lane == 0) PROF_ADD(1); }   // B operand staged, TMEM drained
        act_phase ^= 1;
        tc_fence_after();
        for (int mb = 0; mb < m_blocks; ++mb) {
          const uint32_t d_main = tm + (uint32_t)(mb * 2 * NP);  // columns [0,NP): hi*hi, [NP,2NP): cross terms
          const uint32_t d_cross = d_main + NP;
          if (group == 2 && (k_chunks & 1) == 0) {
            // two weight stages per loop iteration (a CTA that streams several 64-point tiles: 1.68 -> 1.57 ms for
            // the 64 000-point forward+gradient sweep)
            for (int kc = 0; kc < k_chunks; kc += 2) {
              const uint32_t st0 = stage, st1 = stage + 1 == NSTAGE ? 0u : stage + 1;
              const uint32_t ph1 = stage + 1 == NSTAGE ? phase ^ 1u : phase;
              { PROF_T0(); mbar_wait(bar_full + 8 * st0, phase); mbar_wait(bar_full + 8 * st1, ph1); if (lane == 0) PROF_ADD(2); }
              tc_fence_after();
              if (elect_one()) {
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                  const uint32_t st = i ? st1 : st0;
                  const uint64_t da_hi = desc_a_base + (uint64_t)((st * TILE_BYTES) >> 4);
                  const uint64_t da_lo = da_hi + (uint64_t)(TILE_HALF_BYTES >> 4);
                  const uint64_t db = desc_b_base + (uint64_t)(((kc + i) * (KC / 8) * BCH) >> 4);
#pragma unroll
                  for (int j = 0; j < KC / 16; ++j) {
                    umma_f16(d_main, da_hi + (uint64_t)((j * 2 * A_LBO) >> 4), db + (uint64_t)((j * 2 * BCH) >> 4),
                             kIdMain, (kc | i | j) ? 1u : 0u);
                    umma_f16(d_cross, da_lo + (uint64_t)((j * 2 * A_LBO) >> 4), db + (uint64_t)((j * 2 * BCH) >> 4),
                             kIdCross, 1u);
                  }
                  umma_commit(bar_empty + 8 * st);
                }
              }
              __syncwarp();
              stage = st1 + 1 == NSTAGE ? 0u : st1 + 1;
              if (stage <= st0) phase ^= 1;
            }
            continue;
          }
          if (group == 4 && (k_chunks & 3) == 0) {
            constexpr int grp = 4;
            // four weight stages per loop iteration (16-point tiles, 11-stage ring): the fixed cost of an iteration
            // (barrier wait, fence, election, warp re-convergence: ~170 cycles, tools/umma_issue_bench.cu) is paid once
            // per 16 MMAs - 210 -> 173 us for the band pass of one detection
            for (int kc = 0; kc < k_chunks; kc += grp) {
              uint32_t st[4];
              {
                PROF_T0();
                uint32_t sg = stage, pg = phase;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  if (i < grp) {
                    st[i] = sg;
                    mbar_wait(bar_full + 8 * sg, pg);
                    if (++sg == NSTAGE) { sg = 0; pg ^= 1; }
                  }
                }
                stage = sg; phase = pg;
                if (lane == 0) PROF_ADD(2);
              }
              tc_fence_after();
              if (elect_one()) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  if (i < grp) {
                    const uint64_t da_hi = desc_a_base + (uint64_t)((st[i] * TILE_BYTES) >> 4);
                    const uint64_t da_lo = da_hi + (uint64_t)(TILE_HALF_BYTES >> 4);
                    const uint64_t db = desc_b_base + (uint64_t)(((kc + i) * (KC / 8) * BCH) >> 4);
#pragma unroll
                    for (int j = 0; j < KC / 16; ++j) {
                      umma_f16(d_main, da_hi + (uint64_t)((j * 2 * A_LBO) >> 4), db + (uint64_t)((j * 2 * BCH) >> 4),
                               kIdMain, (kc | i | j) ? 1u : 0u);
                      umma_f16(d_cross, da_lo + (uint64_t)((j * 2 * A_LBO) >> 4), db + (uint64_t)((j * 2 * BCH) >> 4),
                               kIdCross, 1u);
                    }
                    umma_commit(bar_empty + 8 * st[i]);
                  }
                }
              }
              __syncwarp();
            }
            continue;
          }
          for (int kc = 0; kc < k_chunks; ++kc) {
            { PROF_T0(); mbar_wait(bar_full + 8 * stage, phase); if (lane == 0) PROF_ADD(2); }
            tc_fence_after();
            if (elect_one()) {
              const uint64_t da_hi = desc_a_base + (uint64_t)((stage * TILE_BYTES) >> 4);
              const uint64_t da_lo = da_hi + (uint64_t)(TILE_HALF_BYTES >> 4);
              const uint64_t db = desc_b_base + (uint64_t)((kc * (KC / 8) * BCH) >> 4);
#pragma unroll
              for (int j = 0; j < KC / 16; ++j) {
                umma_f16(d_main, da_hi + (uint64_t)((j * 2 * A_LBO) >> 4), db + (uint64_t)((j * 2 * BCH) >> 4),
                         kIdMain, (kc | j) ? 1u : 0u);                                      // W_hi x [H_hi ; H_lo]
                umma_f16(d_cross, da_lo + (uint64_t)((j * 2 * A_LBO) >> 4), db + (uint64_t)((j * 2 * BCH) >> 4),
                         kIdCross, 1u);                                                      // W_lo x H_hi
              }
              umma_commit(bar_empty + 8 * stage);     // frees the weight stage
            }
            __syncwarp();
            if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
          }
        }
        if (elect_one()) umma_commit(bar_acc);         // accumulators of this pass are complete
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int q = warp & 3, ph = warp >> 2;            // TMEM lane quadrant, point half (PT points)
    const int t = q * 32 + lane;                       // 0..127 = TMEM lane = feature within the M block
    const int et = tid;                                // 0..255 index among the epilogue threads
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    uint32_t* masks32 = reinterpret_cast<uint32_t*>(masks);   // [layer][feature][point half]
    uint32_t acc_phase = 0;
    float amax = 0.f;
    for (long long it = 0; it < my_tiles; ++it) {
      const long long pt = it * gridDim.x + blockIdx.x;
      const long long base = pt * NP;
      // ---- stage inputs: inp[c][n] ----
      for (int i = et; i < in_pad * NP; i += NEPI) {
        const int c = i / NP, n = i - c * NP;
        const long long gi = base + n;
        float v = 0.f;
        if (gi < n_rows && c < in0) {
          const long long src = in.index ? (long long)in.index[gi] : gi;
          if (in.inputs) {
            v = in.inputs[src * in0 + c];
          } else {
            const long long b = src / in.points_per_batch, k = src - b * in.points_per_batch;
            if (c < latent) {
              v = in.latent_unit[b * latent + c];
            } else {
              float x, y, z;
              lattice_point(in.lattice, k, x, y, z);
              v = (c - latent) == 0 ? x : (c - latent) == 1 ? y : z;
            }
          }
        }
        inp[i] = v;
        dinp[i] = 0.f;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      // B operand of layer 0: k = input column (padded to one 32-k chunk); thread = (k, point half)
      if (q == 0) {
        const int k = lane;
        unsigned char* row = bop + (k >> 3) * BCH + (k & 7) * 16;
#pragma unroll
        for (int g = 0; g < PT / 8; ++g) {
          float h[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) h[e] = k < in0 ? inp[k * NP + ph * PT + g * 8 + e] * ACT_SCALE : 0.f;
          amax = fmaxf(amax, pack8_store_t<LO>(row, ph * (PT / 8) + g, h));
        }
      }
      fence_async_smem();
      tc_fence_before();
      mbar_arrive(bar_act);

      for (int p = 0; p < npass; ++p) {
        const TcPassDev Ps = T.pass[p];               // by value: keeps the fields in registers
        { PROF_T0(); mbar_wait(bar_acc, acc_phase); if (et == 0) PROF_ADD(3); }
        acc_phase ^= 1;
        tc_fence_after();
        PROF_T0();
        if (Ps.kind == 1) {
          // ---- last Linear: row 0 holds the pre-activation of the sdf ----
          if (q == 0) {                                // whole warp issues the aligned loads; lane 0 owns row 0
            const float bias0 = __ldg(Ps.bias);
#pragma unroll
            for (int g = 0; g < G; ++g) {
              uint32_t vm[GW], vc[GW];
              tmem_ldg(lane_base + ph * PT + g * GW, vm);
              tmem_ldg(lane_base + NP + ph * PT + g * GW, vc);
              tmem_ld_wait();
              if (lane == 0) {
#pragma unroll
                for (int qq = 0; qq < GW; ++qq) {
                  const int n = ph * PT + g * GW + qq;
                  float y = (__uint_as_float(vm[qq]) + __uint_as_float(vc[qq])) * Ps.inv_scale + bias0, gg = 1.f;
                  if (use_tanh) { y = tanhf(y); gg *= 1.f - y * y; }
                  y = tanhf(y);
                  gg *= 1.f - y * y;
                  if (base + n < n_rows) sdf_out[base + n] = y;
                  gbuf[n] = gg;
                }
              }
              __syncwarp();
            }
          }
          tc_fence_before();
          asm volatile("bar.sync 1, 256;" ::: "memory");
          if (want_grad) {
            // backward of the last Linear is an outer product: delta[f][n] = W_last[f] * g[n] * mask[f][n]
            const int hidden = T.last_k;
            for (int mb = 0; mb < 4; ++mb) {
              const int f = mb * 128 + t;
              const float w = f < hidden ? __ldg(T.last_w + f) * BWD_SCALE : 0.f;
              const uint32_t mk = f < hidden ? masks32[((size_t)(num_layers - 2) * 512 + f) * 2 + ph] : 0u;
              unsigned char* row = bop + (f >> 3) * BCH + (f & 7) * 16;
#pragma unroll
              for (int g = 0; g < PT / 8; ++g) {
                float h[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) h[e] = ((mk >> (g * 8 + e)) & 1u) ? w * gbuf[ph * PT + g * 8 + e] : 0.f;
                amax = fmaxf(amax, pack8_store_t<LO>(row, ph * (PT / 8) + g, h));
              }
            }
            fence_async_smem();
            tc_fence_before();
            mbar_arrive(bar_act);
          }
          continue;
        }
        const bool fwd = Ps.kind == 0;
        const int split = fwd ? Ps.rows : (Ps.kind == 2 ? Ps.prev_rows : in0);   // rows below: regular outputs
        for (int mb = 0; mb < Ps.m_blocks; ++mb) {
          const int f = mb * 128 + t;
          const int cls = f < split ? 0 : (f < split + Ps.cat_dim ? 1 : 2);
          const uint32_t tb = lane_base + (uint32_t)(mb * 2 * NP);
          unsigned char* row = bop + (f >> 3) * BCH + (f & 7) * 16;
          const float bias = (fwd && cls == 0) ? __ldg(Ps.bias + f) : 0.f;
          const uint32_t pmask = (Ps.kind == 2 && cls == 0) ? masks32[((size_t)(Ps.layer - 1) * 512 + f) * 2 + ph] : 0u;
          const int cat_row = (Ps.cat_off + f - split) * NP + ph * PT;
          uint32_t mk = 0u;
#pragma unroll
          for (int g = 0; g < G; ++g) {
            uint32_t vm[GW], vc[GW];
            tmem_ldg(tb + ph * PT + g * GW, vm);
            tmem_ldg(tb + NP + ph * PT + g * GW, vc);
            tmem_ld_wait();
            float x[GW];
#pragma unroll
            for (int qq = 0; qq < GW; ++qq)
              x[qq] = (__uint_as_float(vm[qq]) + __uint_as_float(vc[qq])) * Ps.inv_scale;
            if (Ps.kind == 3) {                                    // gradient with respect to the input row
              if (cls == 0) {
#pragma unroll
                for (int qq = 0; qq < GW; ++qq) dinp[f * NP + ph * PT + g * GW + qq] += x[qq];
              }
              continue;
            }
            float h[GW];
            if (fwd) {
              if (cls == 0) {
#pragma unroll
                for (int qq = 0; qq < GW; ++qq) {
                  const float y = x[qq] + bias;
                  const bool on = y > 0.f;
                  if (on) mk |= 1u << (g * GW + qq);
                  h[qq] = on ? y * Ps.out_scale : 0.f;
                }
              } else if (cls == 1) {                               // cat[x, input] feeds the next Linear
#pragma unroll
                for (int qq = 0; qq < GW; ++qq) h[qq] = inp[cat_row + g * GW + qq] * Ps.out_scale;
              } else {
#pragma unroll
                for (int qq = 0; qq < GW; ++qq) h[qq] = 0.f;
              }
            } else {
              if (cls == 0) {
#pragma unroll
                for (int qq = 0; qq < GW; ++qq) h[qq] = ((pmask >> (g * GW + qq)) & 1u) ? x[qq] * Ps.out_scale : 0.f;
              } else {
                if (cls == 1) {                                    // gradient of the concatenated input columns
#pragma unroll
                  for (int qq = 0; qq < GW; ++qq) dinp[cat_row + g * GW + qq] += x[qq];
                }
#pragma unroll
                for (int qq = 0; qq < GW; ++qq) h[qq] = 0.f;
              }
            }
#pragma unroll
            for (int pk = 0; pk < PG; ++pk) {
              float h8[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) h8[e] = h[pk * 8 + e];
              const int pgi = ph * (PT / 8) + g * PG + pk;
              amax = fmaxf(amax, pack8_store_t<LO>(row, pgi, h8));
            }
          }
          if (fwd && want_grad) masks32[((size_t)Ps.layer * 512 + f) * 2 + ph] = mk;   // only the backward reads them
        }
        if (et == 0) PROF_ADD(4 + (Ps.kind == 0 ? 0 : Ps.kind == 2 ? 1 : 2));
        if (Ps.kind == 3) {
          tc_fence_before();
          asm volatile("bar.sync 1, 256;" ::: "memory");
          for (int i = et; i < NP * in0; i += NEPI) {
            const int n = i / in0, c = i - n * in0;
            if (base + n < n_rows) dinput_out[(base + n) * in0 + c] = dinp[c * NP + n];
          }
        } else {
          fence_async_smem();
          tc_fence_before();
          mbar_arrive(bar_act);
        }
      }
      // the next point tile re-stages inp/dinp: make sure everyone is done with them
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }
    if (amax > 60000.f) atomicOr(overflow_flag, 1);   // a scaled operand left the fp16 range
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem_base, TCOLS);
}

// ---------------------------------------------------------------------------------------------
// Coarse lattice pass, wide-tile version.
//
// Measured on B200 (tools/umma_issue_bench.cu, profiles/r01_issue_loop.md): the issuing warp spends
// ~170 cycles per iteration of the weight-stage ring (mbarrier wait, fence, commit) whatever it issues,
// and a tcgen05.mma with M=128, K=16 takes max(N/2, ~52) cycles.  Two N=64 MMAs per stage therefore run
// at 90 cycles per MMA (36 % of the tensor peak), four N=128 MMAs per stage at 65 (98 %).  So this
// kernel makes the point tile as wide as one CTA can hold - up to 128 points (N = 128, B operand
// 128 KB, the four M-block accumulators fill the 512 TMEM columns) - and a weight stage two k-chunks
// deep (four MMAs per ring iteration).
//
// That leaves no room for a second resident tile, so the epilogue (16 warps) is overlapped with the MMAs at
// M-block granularity:
//   * accumulator mb is committed on its own mbarrier; M blocks 0 and 1 are turned into packed fp16
//     rows held in registers while the MMAs of the later M blocks still run (the rows cannot be
//     stored yet: they overwrite the B operand those MMAs read);
//   * once the last MMA of the pass has completed, the rows are stored M block by M block, each
//     followed by its own "rows ready" mbarrier; the next pass starts with its M block 0, whose
//     k-chunks 4q..4q+3 only need the rows of M block q, so its MMAs run under the epilogue of
//     M blocks 2 and 3.
// The tile width is chosen on the host so that the last round of tiles is as full as the others
// (112 points for the 40^3 lattice on 148 SMs).
// ---------------------------------------------------------------------------------------------
constexpr int W_THREADS = 576;           // 16 epilogue warps + weight producer + MMA issuer
constexpr int W_NEPI = 512;

struct WidePlan {
  uint32_t stages, b, inp, bars, tmem_slot, total;
};
template <int W_GROUP>
__host__ __device__ inline WidePlan make_wide_plan(int in0, int npt) {
  constexpr int W_STAGES = W_GROUP == 2 ? 5 : 3, W_STAGE_BYTES = W_GROUP * TILE_HALF_BYTES;
  WidePlan p;
  uint32_t o = 0;
  p.stages = o; o += W_STAGES * W_STAGE_BYTES;
  p.b = o; o += 64 * (uint32_t)npt * 16;                 // 64 chunks of 8 k x (npt/8 point groups x 128 B)
  const uint32_t in_pad = (uint32_t)((in0 + 7) & ~7);
  p.inp = o; o += in_pad * (uint32_t)npt * 4;
  p.bars = o; o += 32 * 8;
  p.tmem_slot = o; o += 16;
  p.total = o;
  return p;
}

template <int W_GROUP>      // weight tiles (k-chunks of 32) per ring stage: 2 -> 4 MMAs per iteration, 4 -> 8
__global__ void __launch_bounds__(W_THREADS, 1)
mlp_tc_coarse_wide_kernel(const TcTable* __restrict__ tabp, const unsigned char* __restrict__ tiles, MlpInputs in,
                          float* __restrict__ sdf_out, int npt) {
  extern __shared__ __align__(1024) unsigned char smem[];
  if (in.count_dev && *in.count_dev <= 0) return;
  const TcTable& T = *tabp;
  const int num_layers = T.num_layers, in0 = T.in0, latent = T.latent, use_tanh = T.use_tanh;
  constexpr int W_STAGES = W_GROUP == 2 ? 5 : 3, W_STAGE_BYTES = W_GROUP * TILE_HALF_BYTES;
  const WidePlan P = make_wide_plan<W_GROUP>(in0, npt);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bars = smem_u32(smem + P.bars);
  const uint32_t bar_full = bars, bar_empty = bars + 8 * W_STAGES, bar_acc = bars + 8 * (2 * W_STAGES),
                 bar_rows = bars + 8 * (2 * W_STAGES + 4);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + P.tmem_slot);
  const int in_pad = (in0 + 7) & ~7;
  const int ch = npt * 16;                 // bytes per 8-k chunk of the B operand
  const int ngroups = npt >> 3;            // 8-point groups per tile
  const long long n_rows = mlp_rows(in);
  const long long num_point_tiles = (n_rows + npt - 1) / npt;

  if (tid == 0) {
    for (int s = 0; s < W_STAGES; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    for (int m = 0; m < 4; ++m) { mbar_init(bar_acc + 8 * m, 1); mbar_init(bar_rows + 8 * m, W_NEPI / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 17) tmem_alloc(smem_u32(smem + P.tmem_slot), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  long long my_tiles = 0;
  for (long long pt = blockIdx.x; pt < num_point_tiles; pt += gridDim.x) ++my_tiles;

  if (warp == 16) {
    // ===================== weight producer =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (long long it = 0; it < my_tiles; ++it) {
        for (int p = 0; p < num_layers; ++p) {
          const int m_blocks = T.pass[p].m_blocks, k_chunks = T.pass[p].k_chunks;
          const unsigned char* src = tiles + T.pass[p].tile0 * TILE_BYTES;
          auto load_stage = [&](int mb, int kc, int cnt) {
            mbar_wait(bar_empty + 8 * stage, phase ^ 1);
            mbar_expect_tx(bar_full + 8 * stage, (uint32_t)cnt * TILE_HALF_BYTES);
            for (int i = 0; i < cnt; ++i)
              bulk_g2s(smem_u32(smem + P.stages + stage * W_STAGE_BYTES + i * TILE_HALF_BYTES),
                       src + (size_t)(mb * k_chunks + kc + i) * TILE_BYTES, TILE_HALF_BYTES, bar_full + 8 * stage);
            if (++stage == W_STAGES) { stage = 0; phase ^= 1; }
          };
          if (W_GROUP == 4 && k_chunks == 16 && m_blocks == 4) {
            // same order as the issuer: M blocks 0 and 1 interleaved by k quarter, then M blocks 2 and 3
            for (int qd = 0; qd < 4; ++qd)
              for (int mb = 0; mb < 2; ++mb) load_stage(mb, qd * 4, 4);
            for (int mb = 2; mb < 4; ++mb)
              for (int qd = 0; qd < 4; ++qd) load_stage(mb, qd * 4, 4);
          } else {
            for (int mb = 0; mb < m_blocks; ++mb)
              for (int kc = 0; kc < k_chunks; kc += W_GROUP) load_stage(mb, kc, min(W_GROUP, k_chunks - kc));
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 17) {
    // ===================== MMA issuer (warp-uniform loop, one elected lane) =====================
    // The loop must stay SHORTER than the MMAs it feeds (4 x N/2 cycles per iteration): a first version with
    // run-time k-chunk counts, predicated groups, 64-bit descriptor arithmetic and debug branches took ~380
    // cycles per iteration of four N=112 MMAs (248 cycles of tensor work) and the tensor pipe idled a third of
    // the time.  Passes with the full 16 k-chunks (every 512-wide layer) therefore take the unrolled path
    // below: compile-time operand offsets, no predicates, descriptor arithmetic on the low word only (the
    // 14-bit start-address field cannot carry: every operand lies below 256 KB); the rows-ready waits only
    // exist in the M block 0 instance.
    uint32_t stage = 0, phase = 0, rows_phase = 0;
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint64_t desc_a_base = make_desc(smem_u32(smem + P.stages), A_LBO, A_SBO);
    const uint64_t desc_b_base = make_desc(smem_u32(smem + P.b), (uint32_t)ch, B_SBO);
    const uint32_t da_hi = (uint32_t)(desc_a_base >> 32), da_lo0 = (uint32_t)desc_a_base;
    const uint32_t db_hi = (uint32_t)(desc_b_base >> 32), db_lo0 = (uint32_t)desc_b_base;
    const uint32_t idesc = idesc_for(npt);
    const uint32_t chq = (uint32_t)(ch >> 4);            // descriptor units per 8-k chunk of B
    auto desc64 = [](uint32_t hi, uint32_t lo) { return ((uint64_t)hi << 32) | (uint64_t)lo; };
    // Everything the unrolled path needs is computed ONCE and pinned in registers (the empty asm keeps the
    // compiler from re-deriving the descriptors from the shared-memory base in every iteration: that
    // re-materialised chain of ~25 dependent uniform-datapath instructions was the ~390 cycles per iteration).
    uint32_t db_it[16 / W_GROUP], db_off[W_GROUP * (KC / 16)];
#pragma unroll
    for (int itq = 0; itq < 16 / W_GROUP; ++itq) {
      db_it[itq] = db_lo0 + (uint32_t)(itq * W_GROUP * (KC / 8)) * chq;
      asm volatile("" : "+r"(db_it[itq]));
    }
#pragma unroll
    for (int i = 0; i < W_GROUP; ++i)
#pragma unroll
      for (int j = 0; j < KC / 16; ++j) {
        db_off[i * (KC / 16) + j] = (uint32_t)(i * (KC / 8) + j * 2) * chq;
        asm volatile("" : "+r"(db_off[i * (KC / 16) + j]));
      }
    uint32_t da_hi_r = da_hi, db_hi_r = db_hi, da_lo0_r = da_lo0, idesc_r = idesc, bar_full_r = bar_full, bar_empty_r = bar_empty;
    asm volatile("" : "+r"(da_hi_r), "+r"(db_hi_r), "+r"(da_lo0_r), "+r"(idesc_r), "+r"(bar_full_r), "+r"(bar_empty_r));
    // one ring iteration: W_GROUP weight tiles x 2 MMAs on accumulator d, B rows starting at descriptor word db_lo
    auto ring_iteration = [&](const uint32_t d, const uint32_t db_lo, const bool first) {
      mbar_wait(bar_full_r + 8 * stage, phase);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t da_lo = da_lo0_r + stage * (uint32_t)(W_STAGE_BYTES >> 4);
#pragma unroll
        for (int i = 0; i < W_GROUP; ++i)
#pragma unroll
          for (int j = 0; j < KC / 16; ++j)
            umma_f16(d, desc64(da_hi_r, da_lo + (uint32_t)((i * TILE_HALF_BYTES + j * 2 * A_LBO) >> 4)),
                     desc64(db_hi_r, db_lo + db_off[i * (KC / 16) + j]), idesc_r, (first && i == 0 && j == 0) ? 0u : 1u);
        umma_commit(bar_empty_r + 8 * stage);
      }
      __syncwarp();
      if (++stage == W_STAGES) { stage = 0; phase ^= 1; }
    };
    auto issue_full_block = [&](const uint32_t d, const bool first_block) {
#pragma unroll
      for (int itq = 0; itq < 16 / W_GROUP; ++itq) {
        constexpr int kGroupsPerQuarter = 4 / W_GROUP;   // iterations per 4 k-chunks (rows of one M block)
        if (first_block && (itq % kGroupsPerQuarter) == 0)
          mbar_wait(bar_rows + 8 * (itq / kGroupsPerQuarter), rows_phase);
        ring_iteration(d, db_it[itq], itq == 0);
      }
    };
    for (long long it = 0; it < my_tiles; ++it) {
      for (int p = 0; p < num_layers; ++p) {
        const int m_blocks = __ldg(&T.pass[p].m_blocks), k_chunks = __ldg(&T.pass[p].k_chunks);
        if (W_GROUP == 4 && k_chunks == 16 && m_blocks == 4) {
          // Full pass.  The k-chunks 4q..4q+3 are the rows M block q of the previous pass produced, and its
          // epilogue publishes them in that order (rows_ready[q]) while this pass already runs: M blocks 0 and 1
          // (whose accumulators the epilogue drained first) advance together, one k quarter at a time, so
          // that 2 x 16 MMAs are issued before the last rows are needed.
#pragma unroll
          for (int qd = 0; qd < 4; ++qd) {
            mbar_wait(bar_rows + 8 * qd, rows_phase);
            ring_iteration(tm, db_it[qd], qd == 0);
            ring_iteration(tm + 128u, db_it[qd], qd == 0);
          }
          rows_phase ^= 1;
          if (elect_one()) { umma_commit(bar_acc); umma_commit(bar_acc + 8); }
          __syncwarp();
          for (int mb = 2; mb < 4; ++mb) {
            issue_full_block(tm + (uint32_t)(mb * 128), false);
            if (elect_one()) umma_commit(bar_acc + 8 * mb);
            __syncwarp();
          }
          continue;
        }
        for (int mb = 0; mb < m_blocks; ++mb) {
          const uint32_t d = tm + (uint32_t)(mb * 128);
          if (k_chunks == 16) {
            if (mb == 0) issue_full_block(d, true); else issue_full_block(d, false);
          } else {
            for (int kc = 0; kc < k_chunks; kc += W_GROUP) {
              // k-chunks 4q..4q+3 are the rows the previous pass's M block q produced (for the first pass: the
              // staged inputs); rows_ready[q] also says that accumulator q of the previous pass has been drained
              if (mb == 0 && (kc & 3) == 0) mbar_wait(bar_rows + 8 * (kc >> 2), rows_phase);
              const int cnt = min(W_GROUP, k_chunks - kc);
              mbar_wait(bar_full + 8 * stage, phase);
              tc_fence_after();
              if (elect_one()) {
                const uint32_t da_lo = da_lo0 + ((stage * W_STAGE_BYTES) >> 4);
                const uint32_t db_lo = db_lo0 + (uint32_t)(kc * (KC / 8)) * chq;
#pragma unroll
                for (int i = 0; i < W_GROUP; ++i) {
                  if (i < cnt) {
#pragma unroll
                    for (int j = 0; j < KC / 16; ++j)
                      umma_f16(d, desc64(da_hi, da_lo + (uint32_t)((i * TILE_HALF_BYTES + j * 2 * A_LBO) >> 4)),
                               desc64(db_hi, db_lo + (uint32_t)(i * (KC / 8) + j * 2) * chq), idesc, (kc | i | j) ? 1u : 0u);
                  }
                }
                umma_commit(bar_empty + 8 * stage);
              }
              __syncwarp();
              if (++stage == W_STAGES) { stage = 0; phase ^= 1; }
            }
          }
          if (mb == 0) {                                   // quarters this pass has no k-chunks for
            for (int qq = (k_chunks + 3) >> 2; qq < 4; ++qq) mbar_wait(bar_rows + 8 * qq, rows_phase);
            rows_phase ^= 1;
          }
          if (elect_one()) umma_commit(bar_acc + 8 * mb);   // this M block's accumulator is complete
          __syncwarp();
        }
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int q = warp & 3, slot = warp >> 2;          // TMEM lane quadrant; point groups g = slot, slot+4, ...
    const int tl = q * 32 + lane;                      // TMEM lane = feature within the M block
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    float* inp = reinterpret_cast<float*>(smem + P.inp);
    unsigned char* bop = smem + P.b;
    uint32_t acc_phase = 0;                            // bit mb = phase of bar_acc[mb]
    for (long long it = 0; it < my_tiles; ++it) {
      const long long pt = it * gridDim.x + blockIdx.x;
      const long long base = pt * npt;
      for (int i = tid; i < in_pad * npt; i += W_NEPI) {
        const int c = i / npt, n = i - c * npt;
        const long long gi = base + n;
        float v = 0.f;
        if (gi < n_rows && c < in0) {
          const long long src = in.index ? (long long)in.index[gi] : gi;
          if (in.inputs) {
            v = in.inputs[src * in0 + c];
          } else {
            const long long b = src / in.points_per_batch, k = src - b * in.points_per_batch;
            if (c < latent) {
              v = in.latent_unit[b * latent + c];
            } else {
              float x, y, z;
              lattice_point(in.lattice, k, x, y, z);
              v = (c - latent) == 0 ? x : (c - latent) == 1 ? y : z;
            }
          }
        }
        inp[i] = v;
      }
      asm volatile("bar.sync 1, 512;" ::: "memory");
      if (q == 0) {                                    // B operand of layer 0: k = input column
        const int k = lane;
        unsigned char* row = bop + (k >> 3) * ch + (k & 7) * 16;
#pragma unroll
        for (int gi = 0; gi < 4; ++gi) {
          const int g = slot + 4 * gi;
          if (g < ngroups) {
            float h[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) h[e] = k < in0 ? inp[k * npt + g * 8 + e] * ACT_SCALE : 0.f;
            pack8_store_hi(row, g, h);
          }
        }
      }
      fence_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
#pragma unroll
        for (int m2 = 0; m2 < 4; ++m2) mbar_arrive(bar_rows + 8 * m2);
      }

      for (int p = 0; p < num_layers; ++p) {
        const TcPassDev Ps = T.pass[p];
        if (Ps.kind == 1) {                            // last Linear: row 0 is the pre-activation of the sdf
          mbar_wait(bar_acc, acc_phase & 1u);
          acc_phase ^= 1u;
          tc_fence_after();
          if (q == 0) {
            const float bias0 = __ldg(Ps.bias);
#pragma unroll
            for (int gi = 0; gi < 4; ++gi) {
              const int g = slot + 4 * gi;
              if (g < ngroups) {
                uint32_t v[8];
                tmem_ld8(lane_base + (uint32_t)(g * 8), v);
                tmem_ld_wait();
                if (lane == 0) {
#pragma unroll
                  for (int e = 0; e < 8; ++e) {
                    const int n = g * 8 + e;
                    float y = __uint_as_float(v[e]) * Ps.inv_scale + bias0;
                    if (use_tanh) y = tanhf(y);
                    y = tanhf(y);
                    if (base + n < n_rows) sdf_out[base + n] = y;
                  }
                }
                __syncwarp();
              }
            }
          }
          tc_fence_before();
          continue;
        }
        // M blocks 0 and 1 become packed rows (held in registers) while the MMAs of the later M blocks run.
        // The scales are powers of two, so max(acc * (inv_scale * out_scale) + bias * out_scale, 0) is bit-identical
        // to relu(acc * inv_scale + bias) * out_scale: two instructions per value instead of four.
        const float k_scale = Ps.inv_scale * Ps.out_scale;
        auto process_block = [&](const int mb, uint4 (&cur)[4]) {
          const int f = mb * 128 + tl;
          // the TMEM loads are warp-collective (.sync.aligned): every lane issues them, whatever its row class
          const bool regular = f < Ps.rows;             // all but the concat / padding rows
          const bool cat = !regular && f < Ps.rows + Ps.cat_dim;
          const float bias_s = regular ? __ldg(Ps.bias + f) * Ps.out_scale : 0.f;
          const int cat_row = (Ps.cat_off + f - Ps.rows) * npt;
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            uint32_t v[2][8];
#pragma unroll
            for (int gj = 0; gj < 2; ++gj)
              if (slot + 4 * (2 * half + gj) < ngroups)
                tmem_ld8(lane_base + (uint32_t)(mb * 128 + (slot + 4 * (2 * half + gj)) * 8), v[gj]);
            tmem_ld_wait();
#pragma unroll
            for (int gj = 0; gj < 2; ++gj) {
              const int g = slot + 4 * (2 * half + gj);
              if (g < ngroups) {
                uint32_t* hw = reinterpret_cast<uint32_t*>(&cur[2 * half + gj]);
                if (regular) {
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    const float h0 = fmaxf(fmaf(__uint_as_float(v[gj][2 * i]), k_scale, bias_s), 0.f);
                    const float h1 = fmaxf(fmaf(__uint_as_float(v[gj][2 * i + 1]), k_scale, bias_s), 0.f);
                    const __half2 a = __floats2half2_rn(h0, h1);
                    hw[i] = *reinterpret_cast<const uint32_t*>(&a);
                  }
                } else {
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    float h0 = 0.f, h1 = 0.f;
                    if (cat) { h0 = inp[cat_row + g * 8 + 2 * i] * Ps.out_scale; h1 = inp[cat_row + g * 8 + 2 * i + 1] * Ps.out_scale; }
                    const __half2 a = __floats2half2_rn(h0, h1);
                    hw[i] = *reinterpret_cast<const uint32_t*>(&a);
                  }
                }
              }
            }
          }
        };
        auto publish_block = [&](const int mb, const uint4 (&rowsv)[4]) {
          const int f = mb * 128 + tl;
          unsigned char* row = bop + (f >> 3) * ch + (f & 7) * 16;
#pragma unroll
          for (int gi = 0; gi < 4; ++gi)
            if (slot + 4 * gi < ngroups) *reinterpret_cast<uint4*>(row + (slot + 4 * gi) * 128) = rowsv[gi];
          fence_async_smem();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_rows + 8 * mb);
        };
        uint4 packed[2][4];
        const int hold = Ps.m_blocks >= 3 ? 2 : 0;
#pragma unroll
        for (int mb = 0; mb < 4; ++mb) {
          if (mb < Ps.m_blocks) {
            if (mb < hold) {
              mbar_wait(bar_acc + 8 * mb, (acc_phase >> mb) & 1u);
            } else if (mb == hold) {                   // wait for every remaining accumulator of the pass
#pragma unroll
              for (int m2 = 0; m2 < 4; ++m2)
                if (m2 >= hold && m2 < Ps.m_blocks) mbar_wait(bar_acc + 8 * m2, (acc_phase >> m2) & 1u);
            }
            acc_phase ^= 1u << mb;
            tc_fence_after();
            if (mb < hold) {
              if (mb < 2) process_block(mb, packed[mb < 2 ? mb : 0]);
            } else {
              // every MMA of this pass has completed: the B operand may be overwritten in place.  Rows are
              // published M block by M block so that the next pass can start on the first ones.
              if (mb == hold) {
#pragma unroll
                for (int m2 = 0; m2 < 2; ++m2)
                  if (m2 < hold) publish_block(m2, packed[m2]);
              }
              uint4 cur[4];
              process_block(mb, cur);
              publish_block(mb, cur);
            }
          }
        }
        if (lane == 0) for (int m2 = Ps.m_blocks; m2 < 4; ++m2) mbar_arrive(bar_rows + 8 * m2);
      }
      // the inputs (concatenated rows) are dead once every epilogue thread is past the last pass
      asm volatile("bar.sync 1, 512;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 17) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------
// Coarse lattice pass, CTA-pair version (tcgen05 cta_group::2).
//
// The wide-tile kernel above is bound by the shared-memory data pipe: every 4 KB of a streamed weight tile
// is written once and read once for a single N = 112 MMA (92 wavefronts per 56 cycles of math,
// profiles/r01_mlp_tc_coarse_wide_kernel_ncu.md).  Here the GEMM is issued the usual way round,
//     D[point, feature] = H[point, k] * W^T[k, feature],
// as ONE instruction over a pair of CTAs: M = 256 points (each CTA keeps the activations of its own 128
// points as the K-major A operand), N = 256 features of which each CTA streams and holds only HALF (its
// 128-feature weight tile is the B operand rows it contributes; the hardware shares both halves), and each
// CTA accumulates [its 128 points x 256 features] in its own TMEM.  Per 128-cycle MMA a CTA's shared memory
// now moves 32 (A) + 32 (own half of B) + 32 (bulk-copy write of that half) wavefronts: 75 % of the pipe.
// Because every CTA produces complete feature rows for its OWN points, the next layer's A operand is
// written locally: the only things crossing the pair are the operand halves inside the MMA, the commit
// multicasts and a few mbarrier arrives.
//
//   warps 0-15  epilogue: TMEM lane quadrant (32 points) x 64-column slice of a 256-feature block
//   warp 16     weight producer (bulk copies of this CTA's tiles: M block 2 j + rank of the pass table)
//   warp 17     rank 0: MMA issuer for the pair; rank 1: relays "my stage is full" to rank 0
//
// Per pass the 512 output features are two 256-column blocks j = 0, 1 of the 512 TMEM columns.  Block j of
// the next pass is issued over k-half 0 first (the features block 0 of this pass produced) while the
// epilogue of block 1 is still running, so only the tail of a tile is exposed.
// ---------------------------------------------------------------------------------------------
constexpr int P_THREADS = 576;
constexpr int P_NEPI = 512;
constexpr int P_STAGES = 5;
constexpr int P_STAGE_TILES = 2;                      // k-chunks (32 k) per ring stage
constexpr int P_STAGE_BYTES = P_STAGE_TILES * TILE_HALF_BYTES;
constexpr int P_PTS = 128;                            // points per CTA
constexpr int P_A_KK = 2048;                          // bytes per 8-k group of the A operand: 16 point groups x 128 B

struct PairPlan {
  uint32_t stages, a, inp, bias, bars, tmem_slot, total;
};
__host__ __device__ inline PairPlan make_pair_plan(int in0) {
  PairPlan p;
  uint32_t o = 0;
  p.stages = o; o += P_STAGES * P_STAGE_BYTES;
  p.a = o; o += 64 * P_A_KK;
  const uint32_t in_pad = (uint32_t)((in0 + 7) & ~7);
  p.inp = o; o += in_pad * P_PTS * 4;
  p.bias = o; o += (P_NEPI / 32) * 128 * 4;           // per epilogue warp: the 2 x 64 scaled biases of its column slices
  p.bars = o; o += 32 * 8;
  p.tmem_slot = o; o += 16;
  p.total = o;
  return p;
}

__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// arrive on an mbarrier of another CTA of the cluster (address from mapa)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t remote_bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(remote_bar) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAITC_%=:\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONEC_%=;\n\t"
      "bra WAITC_%=;\n\t"
      "DONEC_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// D[tmem of both CTAs] (+)= A[smem of both] * B[smem of both], issued by one thread of the pair's rank-0 CTA
__device__ __forceinline__ void umma_f16_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the mbarrier at this offset in every CTA of `mask` once the pair MMAs issued so far have completed
__device__ __forceinline__ void umma_commit_pair(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(mask)
               : "memory");
}
// instruction descriptor: D fp32, A / B fp16, both K-major, M = 256 over the pair
__host__ __device__ constexpr uint32_t idesc_pair(int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}

__global__ void __launch_bounds__(P_THREADS, 1)
mlp_tc_coarse_pair_kernel(const TcTable* __restrict__ tabp, const unsigned char* __restrict__ tiles, MlpInputs in,
                          float* __restrict__ sdf_out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  // march mode: clear the counter the step after the next appends to (nobody touches it during this launch)
  if (in.march && blockIdx.x == 0 && threadIdx.x == 0 && in.march->reset_count) *in.march->reset_count = 0;
  if (in.count_dev && *in.count_dev <= 0) return;     // uniform over the grid: before any cluster traffic
  const TcTable& T = *tabp;
  const int num_layers = T.num_layers, in0 = T.in0, latent = T.latent, use_tanh = T.use_tanh;
  const PairPlan P = make_pair_plan(in0);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const uint32_t bars = smem_u32(smem + P.bars);
  const uint32_t bar_full = bars, bar_peer = bars + 8 * P_STAGES, bar_empty = bars + 8 * (2 * P_STAGES),
                 bar_acc = bars + 8 * (3 * P_STAGES), bar_a = bars + 8 * (3 * P_STAGES + 2);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + P.tmem_slot);
  const int in_pad = (in0 + 7) & ~7;
  // march mode: 2^march_lk sample points per listed ray (trace.cuh)
  const int march_lk = in.march ? march_log2k(*in.march, *in.march->count) : 0;
  const long long n_rows = in.march ? march_rows(*in.march) : mlp_rows(in);
#ifdef SDFR_TRACE_DEBUG
  if (in.march && blockIdx.x == 0 && threadIdx.x == 0)
    printf("march launch: %d rays x %d samples, near so far %d\n", *in.march->count, 1 << march_lk, *in.march->near_count);
#endif
  const long long num_pair_tiles = (n_rows + 2 * P_PTS - 1) / (2 * P_PTS);
  const long long num_pairs = gridDim.x >> 1, pair_id = blockIdx.x >> 1;
  // The last Linear (hidden -> 1) is a dot product per point: when the layer before it is a plain hidden layer
  // its epilogue computes it from the fp32 accumulators (no fp16 rounding of that layer's activations), and the
  // 32 N = 16 MMAs, the operand stores and the barrier round of a separate pass disappear.
  const bool fuse_last = num_layers >= 2 && T.pass[num_layers - 1].kind == 1 && T.pass[num_layers - 2].kind == 0 &&
                         T.pass[num_layers - 2].cat_dim == 0 && T.pass[num_layers - 2].rows == T.last_k &&
                         T.pass[num_layers - 2].m_blocks == 4;
  const int num_passes = fuse_last ? num_layers - 1 : num_layers;

  if (tid == 0) {
    for (int s = 0; s < P_STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_peer + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int j = 0; j < 2; ++j) mbar_init(bar_acc + 8 * j, 1);
    for (int qt = 0; qt < 4; ++qt) mbar_init(bar_a + 8 * qt, 2 * (P_NEPI / 32));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 17) tmem_alloc_pair(smem_u32(smem + P.tmem_slot), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                 // the peer's barriers and TMEM exist before anything targets them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  long long my_tiles = 0;
  for (long long pt = pair_id; pt < num_pair_tiles; pt += num_pairs) ++my_tiles;

  if (warp == 16) {
    // ===================== weight producer (both CTAs) =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (long long it = 0; it < my_tiles; ++it) {
        for (int p = 0; p < num_passes; ++p) {
          const int m_blocks = T.pass[p].m_blocks, k_chunks = T.pass[p].k_chunks;
          const unsigned char* src = tiles + T.pass[p].tile0 * TILE_BYTES;
          const int nblocks = T.pass[p].kind == 1 ? 1 : m_blocks >> 1;
          for (int j = 0; j < nblocks; ++j) {
            // hidden pass: this CTA holds features [256 j + 128 rank, + 128) = M block 2 j + rank of the table;
            // last pass: both CTAs load the single tile row (8 of its rows each enter the N = 16 MMA)
            const int mb = T.pass[p].kind == 1 ? 0 : 2 * j + (int)rank;
            for (int kc = 0; kc < k_chunks; kc += P_STAGE_TILES) {
              const int cnt = min(P_STAGE_TILES, k_chunks - kc);
              mbar_wait(bar_empty + 8 * stage, phase ^ 1);
              mbar_expect_tx(bar_full + 8 * stage, (uint32_t)cnt * TILE_HALF_BYTES);
              for (int i = 0; i < cnt; ++i)
                bulk_g2s(smem_u32(smem + P.stages + stage * P_STAGE_BYTES + i * TILE_HALF_BYTES),
                         src + (size_t)(mb * k_chunks + kc + i) * TILE_BYTES, TILE_HALF_BYTES, bar_full + 8 * stage);
              if (++stage == P_STAGES) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 17 && rank == 1) {
    // ===================== rank 1: tell the issuer that this CTA's half of a stage has landed ==============
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      uint32_t peer_bar[P_STAGES];
#pragma unroll
      for (int s = 0; s < P_STAGES; ++s) peer_bar[s] = mapa_u32(bar_peer + 8 * s, 0u);
      for (long long it = 0; it < my_tiles; ++it) {
        for (int p = 0; p < num_passes; ++p) {
          const int nblocks = T.pass[p].kind == 1 ? 1 : T.pass[p].m_blocks >> 1;
          const int steps = nblocks * ((T.pass[p].k_chunks + P_STAGE_TILES - 1) / P_STAGE_TILES);
          for (int st = 0; st < steps; ++st) {
            mbar_wait(bar_full + 8 * stage, phase);
#pragma unroll
            for (int s = 0; s < P_STAGES; ++s)
              if (s == (int)stage) mbar_arrive_cluster(peer_bar[s]);
            if (++stage == P_STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 17) {
    // ===================== rank 0: MMA issuer for the pair =====================
    uint32_t stage = 0, phase = 0, a_phase = 0;
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint64_t desc_a0 = make_desc(smem_u32(smem + P.a), P_A_KK, 128);        // K-major: LBO = next 8 k, SBO = next 8 points
    const uint64_t desc_b0 = make_desc(smem_u32(smem + P.stages), A_LBO, A_SBO);  // a weight tile, K-major
    constexpr uint32_t kIdHidden = idesc_pair(256), kIdLast = idesc_pair(16);
    for (long long it = 0; it < my_tiles; ++it) {
      for (int p = 0; p < num_passes; ++p) {
        const int m_blocks = __ldg(&T.pass[p].m_blocks), k_chunks = __ldg(&T.pass[p].k_chunks);
        const bool last = __ldg(&T.pass[p].kind) == 1;
        const int nblocks = last ? 1 : m_blocks >> 1;
        const uint32_t idesc = last ? kIdLast : kIdHidden;
        // a_ready[qt]: both CTAs have written features [128 qt, 128 qt + 128) of their A operand = k-chunks
        // 4 qt .. 4 qt + 3 of this pass.  The epilogue reads a whole 256-column TMEM block before it publishes the
        // block's first quarter, so a_ready[0] also says "block 0 is drained" and a_ready[3] "block 1 is drained".
        uint32_t waited = 0;                           // bit qt: a_ready[qt] of this pass has been consumed
        auto need = [&](const int qt) {
          if (!((waited >> qt) & 1u)) {
            mbar_wait(bar_a + 8 * qt, a_phase);
            waited |= 1u << qt;
            tc_fence_after();
          }
        };
        need(0);
        for (int j = 0; j < nblocks; ++j) {
          const uint32_t d = tm + (uint32_t)(j * 256);
          if (j == 1) { need(1); need(2); need(3); }
          for (int kc = 0; kc < k_chunks; kc += P_STAGE_TILES) {
            need(kc >> 2);
            const int cnt = min(P_STAGE_TILES, k_chunks - kc);
            mbar_wait(bar_full + 8 * stage, phase);
            mbar_wait(bar_peer + 8 * stage, phase);
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
              for (int i = 0; i < P_STAGE_TILES; ++i) {
                if (i < cnt) {
#pragma unroll
                  for (int t = 0; t < KC / 16; ++t) {
                    const uint64_t da = desc_a0 + (uint64_t)((((kc + i) * (KC / 8) + 2 * t) * P_A_KK) >> 4);
                    const uint64_t db = desc_b0 + (uint64_t)((stage * P_STAGE_BYTES + i * TILE_HALF_BYTES + t * 2 * A_LBO) >> 4);
                    umma_f16_pair(d, da, db, idesc, (kc | i | t) ? 1u : 0u);
                  }
                }
              }
              umma_commit_pair(bar_empty + 8 * stage, (uint16_t)3);    // the stage is free in both CTAs
            }
            __syncwarp();
            if (++stage == P_STAGES) { stage = 0; phase ^= 1; }
          }
          if (elect_one()) umma_commit_pair(bar_acc + 8 * j, (uint16_t)3);   // block j is complete in both CTAs
          __syncwarp();
        }
        need(1); need(2); need(3);                     // keep the phases of the four barriers in step
        a_phase ^= 1;
      }
    }
  } else {
    // ===================== epilogue warps (both CTAs) =====================
    const int q = warp & 3, cs = warp >> 2;            // TMEM lane quadrant (32 points); 64-column slice of a block
    const int pt_l = q * 32 + lane;                    // point of this thread = TMEM lane
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    float* inp = reinterpret_cast<float*>(smem + P.inp);           // [in_pad][128]
    float* wbias = reinterpret_cast<float*>(smem + P.bias) + warp * 128;
    unsigned char* aop = smem + P.a;
    unsigned char* arow = aop + (pt_l >> 3) * 128 + (pt_l & 7) * 16;   // + kk * P_A_KK
    uint32_t a_ready[4];                               // in the issuer's CTA
#pragma unroll
    for (int qt = 0; qt < 4; ++qt) a_ready[qt] = mapa_u32(bar_a + 8 * qt, 0u);
    uint32_t acc_phase = 0;                            // bit j = phase of bar_acc[j]
    auto publish = [&](const uint32_t remote_bar) {
      fence_async_smem();                              // generic-proxy stores -> tensor-core reads
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(remote_bar);
    };
    for (long long it = 0; it < my_tiles; ++it) {
      const long long ptile = it * num_pairs + pair_id;
      const long long base = ptile * (2 * P_PTS) + (long long)rank * P_PTS;
      for (int i = tid; i < in_pad * P_PTS; i += P_NEPI) {
        const int c = i / P_PTS, n = i - c * P_PTS;
        const long long gi = base + n;
        float v = 0.f;
        if (gi < n_rows && c < in0 && in.march) {
          v = march_input(*in.march, gi, c, march_lk);
        } else if (gi < n_rows && c < in0) {
          const long long src = in.index ? (long long)in.index[gi] : gi;
          if (in.inputs) {
            v = in.inputs[src * in0 + c];
          } else {
            const long long b = src / in.points_per_batch, k = src - b * in.points_per_batch;
            if (c < latent) {
              v = in.latent_unit[b * latent + c];
            } else {
              float x, y, z;
              lattice_point(in.lattice, k, x, y, z);
              v = (c - latent) == 0 ? x : (c - latent) == 1 ? y : z;
            }
          }
        }
        inp[i] = v;
      }
      asm volatile("bar.sync 1, 512;" ::: "memory");
      if (cs == 0) {                                   // A operand of layer 0: k = input column, one 32-k chunk
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          uint4 pk;
          uint32_t* hw = reinterpret_cast<uint32_t*>(&pk);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int k0 = kk * 8 + 2 * e;
            const float h0 = k0 < in0 ? inp[k0 * P_PTS + pt_l] * ACT_SCALE : 0.f;
            const float h1 = k0 + 1 < in0 ? inp[(k0 + 1) * P_PTS + pt_l] * ACT_SCALE : 0.f;
            const __half2 a = __floats2half2_rn(h0, h1);
            hw[e] = *reinterpret_cast<const uint32_t*>(&a);
          }
          *reinterpret_cast<uint4*>(arow + kk * P_A_KK) = pk;
        }
      }
      publish(a_ready[0]);
      if (lane == 0) { mbar_arrive_cluster(a_ready[1]); mbar_arrive_cluster(a_ready[2]); mbar_arrive_cluster(a_ready[3]); }

      for (int p = 0; p < num_passes; ++p) {
        const TcPassDev Ps = T.pass[p];
        if (Ps.kind == 1) {                            // last Linear: column 0 is the pre-activation of the sdf
          mbar_wait(bar_acc, acc_phase & 1u);
          acc_phase ^= 1u;
          tc_fence_after();
          if (cs == 0) {
            uint32_t v[8];
            tmem_ld8(lane_base, v);
            tmem_ld_wait();
            float y = __uint_as_float(v[0]) * Ps.inv_scale + __ldg(Ps.bias);
            if (use_tanh) y = tanhf(y);
            y = tanhf(y);
            if (in.march) march_advance(*in.march, base + pt_l < n_rows, base + pt_l, y, march_lk);
            else if (base + pt_l < n_rows) sdf_out[base + pt_l] = y;
          }
          tc_fence_before();
          continue;
        }
        if (fuse_last && p == num_passes - 1) {
          // ---- last hidden layer + last Linear: sdf = tanh(sum_f w_f relu(acc_f / scale + b_f) + b_last) ----
          __syncwarp();
#pragma unroll
          for (int qt = 0; qt < 4; ++qt) {
            const int f = qt * 128 + cs * 32 + lane;
            wbias[qt * 32 + lane] = f < Ps.rows ? __ldg(Ps.bias + f) : 0.f;
          }
          __syncwarp();
          float partial = 0.f;
          auto dot_quarter = [&](const int qt) {
            const int f0 = qt * 128 + cs * 32;
            uint32_t v[32];
            tmem_ld16(lane_base + (uint32_t)f0, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
            tmem_ld16(lane_base + (uint32_t)(f0 + 16), *reinterpret_cast<uint32_t(*)[16]>(&v[16]));
            tmem_ld_wait();
            const float* wb = wbias + qt * 32;
#pragma unroll
            for (int o = 0; o < 8; ++o) {
              const float4 w4 = __ldg(reinterpret_cast<const float4*>(T.last_w + f0) + o);
              const float4 b4 = *reinterpret_cast<const float4*>(wb + o * 4);
              partial = fmaf(w4.x, fmaxf(fmaf(__uint_as_float(v[4 * o]), Ps.inv_scale, b4.x), 0.f), partial);
              partial = fmaf(w4.y, fmaxf(fmaf(__uint_as_float(v[4 * o + 1]), Ps.inv_scale, b4.y), 0.f), partial);
              partial = fmaf(w4.z, fmaxf(fmaf(__uint_as_float(v[4 * o + 2]), Ps.inv_scale, b4.z), 0.f), partial);
              partial = fmaf(w4.w, fmaxf(fmaf(__uint_as_float(v[4 * o + 3]), Ps.inv_scale, b4.w), 0.f), partial);
            }
          };
          mbar_wait(bar_acc, acc_phase & 1u);
          acc_phase ^= 1u;
          tc_fence_after();
          dot_quarter(0);
          dot_quarter(1);
          mbar_wait(bar_acc + 8, (acc_phase >> 1) & 1u);   // also: every MMA of the tile has completed
          acc_phase ^= 2u;
          tc_fence_after();
          dot_quarter(2);
          dot_quarter(3);
          tc_fence_before();
          // the inputs are dead by now (last used by the concat layer): their buffer carries the 4 partial sums
          inp[cs * P_PTS + pt_l] = partial;
          asm volatile("bar.sync 1, 512;" ::: "memory");
          if (cs == 0) {
            float y = ((inp[pt_l] + inp[P_PTS + pt_l]) + (inp[2 * P_PTS + pt_l] + inp[3 * P_PTS + pt_l])) +
                      __ldg(T.pass[num_layers - 1].bias);
            if (use_tanh) y = tanhf(y);
            y = tanhf(y);
            if (in.march) march_advance(*in.march, base + pt_l < n_rows, base + pt_l, y, march_lk);
            else if (base + pt_l < n_rows) sdf_out[base + pt_l] = y;
          }
          continue;
        }
        const float k_scale = Ps.inv_scale * Ps.out_scale;
        const int nblocks = Ps.m_blocks >> 1;
        // this warp's scaled biases (features 128 qt + 32 cs + 0..31 of the four quarters), fetched under the MMAs
        __syncwarp();
#pragma unroll
        for (int qt = 0; qt < 4; ++qt) {
          const int f = qt * 128 + cs * 32 + lane;
          wbias[qt * 32 + lane] = f < Ps.rows ? __ldg(Ps.bias + f) * Ps.out_scale : 0.f;
        }
        __syncwarp();
        // quarter qt of the pass output = features [128 qt, +128); this warp's 32-column slice of it for its
        // 32 points -> four packed 16-byte rows of the next A operand per thread
        auto compute_quarter = [&](const int qt, uint4 (&pk)[4]) {
          const int f0 = qt * 128 + cs * 32;
          uint32_t v[32];
          tmem_ld16(lane_base + (uint32_t)f0, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
          tmem_ld16(lane_base + (uint32_t)(f0 + 16), *reinterpret_cast<uint32_t(*)[16]>(&v[16]));
          tmem_ld_wait();
          const float* wb = wbias + qt * 32;
          if (f0 + 32 <= Ps.rows) {                    // all regular rows (warp-uniform)
#pragma unroll
            for (int o = 0; o < 4; ++o) {
              const float4 b0 = *reinterpret_cast<const float4*>(wb + o * 8);
              const float4 b1 = *reinterpret_cast<const float4*>(wb + o * 8 + 4);
              const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
              uint32_t* hw = reinterpret_cast<uint32_t*>(&pk[o]);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float h0 = fmaxf(fmaf(__uint_as_float(v[o * 8 + 2 * e]), k_scale, bb[2 * e]), 0.f);
                const float h1 = fmaxf(fmaf(__uint_as_float(v[o * 8 + 2 * e + 1]), k_scale, bb[2 * e + 1]), 0.f);
                const __half2 a = __floats2half2_rn(h0, h1);
                hw[e] = *reinterpret_cast<const uint32_t*>(&a);
              }
            }
          } else {                                     // the slice that holds the concatenated / padding rows
#pragma unroll
            for (int o = 0; o < 4; ++o) {
              uint32_t* hw = reinterpret_cast<uint32_t*>(&pk[o]);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float h[2];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                  const int f = f0 + o * 8 + 2 * e + u;
                  float val;
                  if (f < Ps.rows) {
                    val = fmaxf(fmaf(__uint_as_float(v[o * 8 + 2 * e + u]), k_scale, wb[o * 8 + 2 * e + u]), 0.f);
                  } else if (f < Ps.rows + Ps.cat_dim) {           // cat[x, input] feeds the next Linear
                    val = inp[(Ps.cat_off + f - Ps.rows) * P_PTS + pt_l] * Ps.out_scale;
                  } else {
                    val = 0.f;
                  }
                  h[u] = val;
                }
                const __half2 a = __floats2half2_rn(h[0], h[1]);
                hw[e] = *reinterpret_cast<const uint32_t*>(&a);
              }
            }
          }
        };
        auto store_quarter = [&](const int qt, const uint4 (&pk)[4]) {
          const int kk0 = (qt * 128 + cs * 32) >> 3;
#pragma unroll
          for (int r = 0; r < 4; ++r) *reinterpret_cast<uint4*>(arow + (kk0 + r) * P_A_KK) = pk[r];
        };
        uint4 pk0[4], pk1[4];
        mbar_wait(bar_acc, acc_phase & 1u);
        acc_phase ^= 1u;
        tc_fence_after();
        compute_quarter(0, pk0);                       // under the MMAs of block 1, which still read all of A
        compute_quarter(1, pk1);
        if (nblocks == 2) {
          mbar_wait(bar_acc + 8, (acc_phase >> 1) & 1u);   // every MMA of the pass has completed: A may be overwritten
          acc_phase ^= 2u;
          tc_fence_after();
        }
        store_quarter(0, pk0);
        publish(a_ready[0]);
        store_quarter(1, pk1);
        publish(a_ready[1]);
        if (nblocks == 2) {
          compute_quarter(2, pk0);
          store_quarter(2, pk0);
          publish(a_ready[2]);
          compute_quarter(3, pk1);
          store_quarter(3, pk1);
          publish(a_ready[3]);
        } else {
          publish(a_ready[2]);
          publish(a_ready[3]);
        }
      }
      // inputs / A operand are rewritten by the next tile: every epilogue thread must be past the last pass
      asm volatile("bar.sync 1, 512;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                 // no CTA may exit (or free TMEM) while its peer can still reach it
  if (warp == 17) tmem_dealloc_pair(tmem_base, TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------
// Accurate forward + input gradient on CTA pairs (tcgen05 cta_group::2): the band pass.
//
// A short row list (the ~1 850 pre-selected lattice points of one detection) is spread over the machine in
// small point tiles, and with 16 points per CTA every tcgen05.mma of mlp_tc_kernel<16> has N <= 32 and sits
// on the ~52-cycle floor of an M = 128 instruction: each CTA sweeps all 932 weight tiles with four such MMAs
// per tile (profiles/r01_launches.md).  Here one instruction covers M = 256 FEATURES over a pair of CTAs:
//   A  = weight tiles, K-major; CTA `rank` streams M block 2 j + rank of the pass table (half of the tiles),
//   B  = activations of the pair's 2 x NP points, MN-major; each CTA holds its OWN NP points, laid out per
//        8-k chunk as [zeros | hi | lo] so that one descriptor starting at `hi` gives the N-rows [hi ; lo] and
//        one starting at `zeros` gives [0 ; hi]:
//            main  MMA  W_hi x [H_hi ; H_lo]   -> columns [hh | hl] of each CTA's half of D
//            cross MMA  W_lo x [0 ; H_hi]      -> adds lo*hi onto the hl columns (and exact zeros onto hh),
//        the same two accumulation chains per output as mlp_tc_kernel, so the results are bit-identical,
//   D  = [128 features of this CTA] x [hh, cross of CTA 0's points | hh, cross of CTA 1's points] in TMEM.
// Two MMAs per K step and 256 features instead of two per 128: half the instructions and half the weight
// stream per CTA.  The epilogue thread that owns a feature writes the next layer's B rows for BOTH point
// halves: its own CTA's with plain shared-memory stores, the peer's through distributed shared memory
// (st.shared::cluster; 16 KB per CTA and pass at NP = 16).
//
//   warps 0-7  epilogue: TMEM lane quadrant (warp & 3) x point half (warp >> 2 = the CTA that owns the points)
//   warp 8     weight producer (both CTAs, own tiles)
//   warp 9     rank 0: MMA issuer for the pair; rank 1: relays "my stage has landed" to rank 0
// ---------------------------------------------------------------------------------------------
constexpr int Q_THREADS = 320;
constexpr int Q_NEPI = 256;
// NP = points per CTA.  16: short row lists; the B operand is laid out [zeros | hi | lo] and a K step is TWO MMAs
// (N = 4 NP).  64: long row lists (many detections per launch); the [hi | lo] layout without the zero groups is what
// fits next to the ring, and a K step is THREE MMAs of N = 2 NP = 128 (W_hi x H_hi -> hh; W_hi x H_lo, W_lo x H_hi -> cross).
// Either way every output sees the same two accumulation chains, in the same order, as in mlp_tc_kernel.
template <int NP> struct BandPad { static constexpr bool value = NP <= 32; };
// BandPipe double-buffers the B operand, so that the epilogue of a 256-feature block runs under the MMAs of the next
// block / the next pass.  Implemented and bit-identical, but measured SLOWER for NP = 16 (115 us against 106 us):
// the kernel is bound by the L2 -> SM weight stream (57 pairs x 14.9 MB in ~100 us is ~8 TB/s of L2 reads), and the
// 96 KB of a second operand come out of the weight ring (112 KB in flight instead of 160 KB).  So it stays off, and
// what remains of it is the half-granular hand-over: the MMAs of the next pass start on k-chunks 0-7 while the
// epilogue is still writing the rows of k-chunks 8-15.
template <int NP> struct BandPipe { static constexpr bool value = false; };
template <int NP> struct BandStageTiles { static constexpr int value = BandPipe<NP>::value ? 1 : (NP <= 32 ? 2 : 1); };   // k-chunks (32 k) per ring stage
template <int NP> struct BandRing { static constexpr int value = BandPipe<NP>::value ? 7 : (NP <= 16 ? 5 : (NP <= 32 ? 3 : 5)); };

struct BandPairPlan {
  uint32_t stages, b, inp, dinp, g, bars, tmem_slot, total;
};
template <int NP>
__host__ __device__ inline BandPairPlan make_band_pair_plan(int in0) {
  BandPairPlan p;
  uint32_t o = 0;
  p.stages = o; o += BandRing<NP>::value * BandStageTiles<NP>::value * TILE_BYTES;
  p.b = o; o += (BandPipe<NP>::value ? 2 : 1) * 64 * ((BandPad<NP>::value ? 3 : 2) * NP * 16);   // 64 8-k chunks x [zeros |] hi | lo
  const uint32_t in_pad = (uint32_t)((in0 + 7) & ~7);
  p.inp = o; o += in_pad * 2 * NP * 4;                 // inputs of the whole pair tile
  p.dinp = o; o += 2 * in_pad * 2 * NP * 4;            // input-gradient partials, double-buffered by tile parity
  p.g = o; o += 2 * NP * 4;
  p.bars = o; o += 32 * 8;
  p.tmem_slot = o; o += 16;
  p.total = o;
  return p;
}

__device__ __forceinline__ void st_cluster_v4(uint32_t addr, const uint4& v) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ float ld_cluster_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void mbar_arrive_cluster_release(uint32_t remote_bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote_bar) : "memory");
}
__device__ __forceinline__ void fence_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// hi / lo split of 8 values -> two 16-byte rows at a shared::cluster address (own or peer CTA)
__device__ __forceinline__ float pack8_store_cluster(uint32_t row_hi, uint32_t lo_off, int pg, const float (&h)[8]) {
  uint4 hi, lo;
  uint32_t* hw = reinterpret_cast<uint32_t*>(&hi);
  uint32_t* lw = reinterpret_cast<uint32_t*>(&lo);
  float amax = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __half2 a = __floats2half2_rn(h[2 * i], h[2 * i + 1]);
    const float2 b = __half22float2(a);
    const __half2 l = __floats2half2_rn(h[2 * i] - b.x, h[2 * i + 1] - b.y);
    hw[i] = *reinterpret_cast<const uint32_t*>(&a);
    lw[i] = *reinterpret_cast<const uint32_t*>(&l);
    amax = fmaxf(amax, fmaxf(fabsf(h[2 * i]), fabsf(h[2 * i + 1])));
  }
  st_cluster_v4(row_hi + pg * 128, hi);
  st_cluster_v4(row_hi + lo_off + pg * 128, lo);
  return amax;
}

// instruction descriptor: D fp32, A / B fp16, A K-major, B MN-major, M = 256 over the pair
__host__ __device__ constexpr uint32_t idesc_band_pair(int n) {
  return (1u << 4) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}

template <int NP>
__global__ void __launch_bounds__(Q_THREADS, 1)
mlp_tc_band_pair_kernel(const TcTable* __restrict__ tabp, const unsigned char* __restrict__ tiles, MlpInputs in,
                        float* __restrict__ sdf_out, float* __restrict__ dinput_out, int* __restrict__ overflow_flag,
                        unsigned long long* __restrict__ mask_scratch) {
  extern __shared__ __align__(1024) unsigned char smem[];
  // uniform over the grid, before any cluster traffic: nothing to evaluate, or a row count outside this launch's window
  if (in.count_dev && (*in.count_dev <= in.count_lo || *in.count_dev > in.count_hi)) return;
  const TcTable& T = *tabp;
  const int num_layers = T.num_layers, in0 = T.in0, latent = T.latent, use_tanh = T.use_tanh;
  constexpr int NSTAGE = BandRing<NP>::value;
  constexpr int Q_STAGE_TILES = BandStageTiles<NP>::value;
  constexpr int Q_STAGE_BYTES = Q_STAGE_TILES * TILE_BYTES;
  constexpr bool kPad = BandPad<NP>::value;
  constexpr bool kPipe = BandPipe<NP>::value;          // two B operands: the epilogue of a block runs under the MMAs of the next
  // hand-over between epilogue and issuer per 256-feature block / k-half (the next pass starts on k-chunks 0-7 while
  // the rows of 8-15 are still being written), or once per pass: the finer protocol costs two more barrier rounds per
  // pass, which pays for the long epilogues of 2 x 64 points (1 246 -> 1 217 us) and not for 2 x 16 (106 -> 112 us)
  constexpr bool kSplit = kPipe || NP > 16;
  constexpr int NPP = 2 * NP;                          // points of the pair tile
  constexpr int BCH = (kPad ? 3 : 2) * NP * 16;        // bytes per 8-k chunk of B: [zeros,] hi, lo point groups
  constexpr int BBYTES = 64 * BCH;                     // one B operand
  constexpr int HI = kPad ? NP * 16 : 0, LO = NP * 16; // hi region inside a chunk; lo region relative to hi
  constexpr int GW = 16;                               // points per TMEM load group
  constexpr int G = NP / GW;
  constexpr int PG = GW / 8;
  constexpr int TCOLS = 8 * NP;                        // two 256-feature blocks x 4 NP columns
  // columns of a block: padded [hh(0) | cross(0) | hh(1) | cross(1)], split [hh(0) | hh(1) | cross(0) | cross(1)]
  constexpr int COL_PH = kPad ? 2 * NP : NP, COL_CROSS = kPad ? NP : 2 * NP;
  constexpr uint32_t kId = idesc_band_pair(kPad ? 4 * NP : 2 * NP);
  static_assert(NP == 16 || NP == 32 || NP == 64, "pair tiles of 2 x 16, 2 x 32 or 2 x 64 points");
  const BandPairPlan P = make_band_pair_plan<NP>(in0);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  unsigned char* bop = smem + P.b;                     // B operand(s): pass g reads operand (g & 1) when kPipe
  unsigned long long* masks = mask_scratch + (size_t)blockIdx.x * (size_t)(num_layers - 1) * 512;
  float* inp = reinterpret_cast<float*>(smem + P.inp);         // [in_pad][2 NP]
  float* gbuf = reinterpret_cast<float*>(smem + P.g);
  const uint32_t bars = smem_u32(smem + P.bars);
  // bar_acc[j]: the MMAs of 256-feature block j of the pass have completed (commit multicast to both CTAs);
  // bar_half[h] (rank 0's copy is used): k-chunks [8 h, 8 h + 8) of the NEXT pass's B operand are written in both
  // CTAs and TMEM block h has been read - one arrival per epilogue warp of both CTAs, every pass
  const uint32_t bar_full = bars, bar_peer = bars + 8 * NSTAGE, bar_empty = bars + 8 * (2 * NSTAGE),
                 bar_acc = bars + 8 * (3 * NSTAGE), bar_half = bars + 8 * (3 * NSTAGE + 2),
                 bar_x = bars + 8 * (3 * NSTAGE + 4);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + P.tmem_slot);
  const int in_pad = (in0 + 7) & ~7;
  const bool want_grad = dinput_out != nullptr;
  const int npass = want_grad ? T.num_passes : num_layers;
  const long long n_rows = mlp_rows(in);
  const long long num_pair_tiles = (n_rows + NPP - 1) / NPP;
  const long long num_pairs = gridDim.x >> 1, pair_id = blockIdx.x >> 1;

  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_peer + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int j = 0; j < 2; ++j) {
      mbar_init(bar_acc + 8 * j, 1);
      mbar_init(bar_half + 8 * j, 2 * (Q_NEPI / 32));
    }
    mbar_init(bar_x, Q_NEPI / 32);                     // the peer's epilogue warps: "my input-gradient partials are final"
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (kPad) {   // the zero point groups of every chunk are written once and never again
    for (int i = tid; i < (kPipe ? 2 : 1) * 64 * NP; i += Q_THREADS)
      *reinterpret_cast<uint4*>(bop + (i / NP) * BCH + (i % NP) * 16) = make_uint4(0u, 0u, 0u, 0u);
    fence_async_smem();
  }
  if (warp == 9) tmem_alloc_pair(smem_u32(smem + P.tmem_slot), TCOLS);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  long long my_tiles = 0;
  for (long long pt = pair_id; pt < num_pair_tiles; pt += num_pairs) ++my_tiles;

  if (warp == 8) {
    // ===================== weight producer (both CTAs) =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (long long it = 0; it < my_tiles; ++it) {
        for (int p = 0; p < npass; ++p) {
          const int m_blocks = T.pass[p].m_blocks, k_chunks = T.pass[p].k_chunks;
          const unsigned char* src = tiles + T.pass[p].tile0 * TILE_BYTES;
          const int nblocks = m_blocks == 1 ? 1 : m_blocks >> 1;
          for (int j = 0; j < nblocks; ++j) {
            // a single-block pass (last Linear, gradient of the first) is computed by both CTAs on the same tile row
            const int mb = m_blocks == 1 ? 0 : 2 * j + (int)rank;
            for (int kc = 0; kc < k_chunks; kc += Q_STAGE_TILES) {
              const int cnt = min(Q_STAGE_TILES, k_chunks - kc);
              mbar_wait(bar_empty + 8 * stage, phase ^ 1);
              mbar_expect_tx(bar_full + 8 * stage, (uint32_t)cnt * TILE_BYTES);
              bulk_g2s(smem_u32(smem + P.stages + stage * Q_STAGE_BYTES), src + (size_t)(mb * k_chunks + kc) * TILE_BYTES,
                       (uint32_t)cnt * TILE_BYTES, bar_full + 8 * stage);
              if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 9 && rank == 1) {
    // ===================== rank 1: tell the issuer that this CTA's half of a stage has landed ==============
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      uint32_t peer_bar[NSTAGE];
#pragma unroll
      for (int s = 0; s < NSTAGE; ++s) peer_bar[s] = mapa_u32(bar_peer + 8 * s, 0u);
      for (long long it = 0; it < my_tiles; ++it) {
        for (int p = 0; p < npass; ++p) {
          const int m_blocks = T.pass[p].m_blocks;
          const int nblocks = m_blocks == 1 ? 1 : m_blocks >> 1;
          const int steps = nblocks * ((T.pass[p].k_chunks + Q_STAGE_TILES - 1) / Q_STAGE_TILES);
          for (int st = 0; st < steps; ++st) {
            mbar_wait(bar_full + 8 * stage, phase);
#pragma unroll
            for (int s = 0; s < NSTAGE; ++s)
              if (s == (int)stage) mbar_arrive_cluster(peer_bar[s]);
            if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 9) {
    // ===================== rank 0: MMA issuer for the pair =====================
    uint32_t stage = 0, phase = 0, half_phase = 0, gpass = 0;
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint64_t desc_a_base = make_desc(smem_u32(smem + P.stages), A_LBO, A_SBO);
    // ring stages per issuer iteration: eight MMAs per iteration either way (the fixed cost of an iteration - barrier
    // waits, fence, election - is ~170 cycles; sixteen per iteration was measured slower: the issuer then waits for the
    // later stage before it starts on the earlier one)
    constexpr int GRP = kPipe ? 2 : 1;
    for (long long it = 0; it < my_tiles; ++it) {
      for (int p = 0; p < npass; ++p, ++gpass) {
        const int m_blocks = __ldg(&T.pass[p].m_blocks), k_chunks = __ldg(&T.pass[p].k_chunks);
        const int nblocks = m_blocks == 1 ? 1 : m_blocks >> 1;
        const uint32_t bsrc = smem_u32(bop) + (kPipe ? (gpass & 1u) * BBYTES : 0u);
        // padded layout: N-rows of a CTA [hi ; lo] (main) and [0 ; hi] (cross); split layout: hi and lo separately
        const uint64_t desc_b_main = make_desc(bsrc + HI, BCH, B_SBO);
        const uint64_t desc_b_cross = make_desc(bsrc, BCH, B_SBO);
        const uint64_t desc_b_lo = make_desc(bsrc + HI + LO, BCH, B_SBO);
        uint32_t waited = 0;                           // bit h: bar_half[h] of this pass has been consumed
        auto need = [&](const int h_) {
          const int h = kSplit ? h_ : 0;
          if (!((waited >> h) & 1u)) {
            mbar_wait_cluster(bar_half + 8 * h, half_phase);
            waited |= 1u << h;
            tc_fence_after();
          }
        };
        need(0);                                       // k-chunks 0-7 staged in both CTAs, TMEM block 0 drained
        for (int j = 0; j < nblocks; ++j) {
          const uint32_t d = tm + (uint32_t)(j * 4 * NP);
          if (j == 1) need(1);                         // TMEM block 1 drained
          for (int kc = 0; kc < k_chunks; kc += Q_STAGE_TILES * GRP) {
            if (kc >= 8) need(1);
            uint32_t st[GRP];
            int nst = 0;
#pragma unroll
            for (int g = 0; g < GRP; ++g) {
              if (kc + g * Q_STAGE_TILES < k_chunks) {
                st[g] = stage;
                mbar_wait(bar_full + 8 * stage, phase);
                mbar_wait(bar_peer + 8 * stage, phase);
                if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                nst = g + 1;
              }
            }
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
              for (int g = 0; g < GRP; ++g) {
                if (g < nst) {
                  const int kc0 = kc + g * Q_STAGE_TILES;
                  const int cnt = min(Q_STAGE_TILES, k_chunks - kc0);
#pragma unroll
                  for (int i = 0; i < Q_STAGE_TILES; ++i) {
                    if (i < cnt) {
                      const uint64_t da_hi = desc_a_base + (uint64_t)((st[g] * Q_STAGE_BYTES + i * TILE_BYTES) >> 4);
                      const uint64_t da_lo = da_hi + (uint64_t)(TILE_HALF_BYTES >> 4);
                      const uint64_t koff = (uint64_t)(((kc0 + i) * (KC / 8) * BCH) >> 4);
#pragma unroll
                      for (int t = 0; t < KC / 16; ++t) {
                        const uint64_t ka = (uint64_t)((t * 2 * A_LBO) >> 4), kb = koff + (uint64_t)((t * 2 * BCH) >> 4);
                        const uint32_t acc = (kc0 | i | t) ? 1u : 0u;
                        if (kPad) {
                          umma_f16_pair(d, da_hi + ka, desc_b_main + kb, kId, acc);                   // W_hi x [H_hi ; H_lo]
                          umma_f16_pair(d, da_lo + ka, desc_b_cross + kb, kId, 1u);                   // W_lo x [0 ; H_hi]
                        } else {
                          umma_f16_pair(d, da_hi + ka, desc_b_main + kb, kId, acc);                   // W_hi x H_hi -> hh
                          umma_f16_pair(d + COL_CROSS, da_hi + ka, desc_b_lo + kb, kId, acc);         // W_hi x H_lo -> cross
                          umma_f16_pair(d + COL_CROSS, da_lo + ka, desc_b_main + kb, kId, 1u);        // W_lo x H_hi -> cross
                        }
                      }
                    }
                  }
                  umma_commit_pair(bar_empty + 8 * st[g], (uint16_t)3);   // the stage is free in both CTAs
                }
              }
            }
            __syncwarp();
          }
          if (kSplit || j == nblocks - 1) {
            if (elect_one()) umma_commit_pair(bar_acc + (kSplit ? 8 * j : 0), (uint16_t)3);   // block j (the pass) is complete in both CTAs
            __syncwarp();
          }
        }
        if (kSplit && nblocks == 1) {                  // keep the second block's barriers in step
          if (elect_one()) umma_commit_pair(bar_acc + 8, (uint16_t)3);
          __syncwarp();
        }
        need(1);
        half_phase ^= 1;
      }
    }
  } else {
    // ===================== epilogue warps (both CTAs) =====================
    const int q = warp & 3, ph = warp >> 2;            // TMEM lane quadrant; point half = CTA that owns the points
    const int t = q * 32 + lane;                       // TMEM lane = feature within this CTA's 128-feature block
    const int et = tid;
    const bool own = ph == (int)rank;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t b_remote = mapa_u32(smem_u32(bop), (uint32_t)ph) + HI;   // hi region of the point owner's first B operand
    const uint32_t half_bar[2] = {mapa_u32(bar_half, 0u), mapa_u32(bar_half + 8, 0u)};
    const uint32_t peer_x = mapa_u32(bar_x, rank ^ 1u);
    // ReLU sign words: [layer][feature of this CTA (block j, lane t)][point half], one u64 each
    auto mask_at = [&](const int layer, const int j) -> unsigned long long& {
      return masks[((size_t)layer * 256 + (size_t)(j * 128 + t)) * 2 + ph];
    };
    uint32_t acc_phase = 0, x_phase = 0, gpass = 0;
    float amax = 0.f;
    auto publish = [&](const int h) {                  // this warp's rows of k-half h (own and remote) are written, TMEM block h read
      if (!kSplit && h == 0) return;                   // one hand-over per pass: after the last block
      fence_async_all();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster_release(half_bar[kSplit ? h : 0]);
    };
    auto wait_acc = [&](const int j_) {
      if (!kSplit && j_ == 0) return;                  // one commit per pass
      const int j = kSplit ? j_ : 0;
      mbar_wait(bar_acc + 8 * j, (acc_phase >> j) & 1u);
      acc_phase ^= 1u << j;
      tc_fence_after();
    };
    for (long long it = 0; it < my_tiles; ++it) {
      const long long ptile = it * num_pairs + pair_id;
      const long long base = ptile * NPP;              // first row of the pair tile; this CTA owns [rank NP, rank NP + NP)
      float* dinp = reinterpret_cast<float*>(smem + P.dinp) + (it & 1) * in_pad * NPP;
      // ---- stage the inputs of the whole pair tile: inp[c][n] ----
      for (int i = et; i < in_pad * NPP; i += Q_NEPI) {
        const int c = i / NPP, n = i - c * NPP;
        const long long gi = base + n;
        float v = 0.f;
        if (gi < n_rows && c < in0) {
          const long long src = in.index ? (long long)in.index[gi] : gi;
          if (in.inputs) {
            v = in.inputs[src * in0 + c];
          } else {
            const long long b = src / in.points_per_batch, k = src - b * in.points_per_batch;
            if (c < latent) {
              v = in.latent_unit[b * latent + c];
            } else {
              float x, y, z;
              lattice_point(in.lattice, k, x, y, z);
              v = (c - latent) == 0 ? x : (c - latent) == 1 ? y : z;
            }
          }
        }
        inp[i] = v;
        dinp[i] = 0.f;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      // B operand of layer 0 (own points): k = input column, one 32-k chunk, into the operand the first pass reads
      if (q == 0 && own) {
        const int k = lane;
        unsigned char* row = bop + (kPipe ? (gpass & 1u) * BBYTES : 0u) + (k >> 3) * BCH + HI + (k & 7) * 16;
#pragma unroll
        for (int g = 0; g < NP / 8; ++g) {
          float h[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) h[e] = k < in0 ? inp[k * NPP + ph * NP + g * 8 + e] * ACT_SCALE : 0.f;
          amax = fmaxf(amax, pack8_store_t<LO>(row, g, h));
        }
      }
      publish(0);
      publish(1);

      for (int p = 0; p < npass; ++p, ++gpass) {
        const TcPassDev Ps = T.pass[p];
        // rows of the next pass's operand: the other buffer when double-buffered, in place otherwise (then only
        // once every MMA of this pass has completed)
        const uint32_t b_dst = b_remote + (kPipe ? ((gpass + 1u) & 1u) * BBYTES : 0u);
        if (Ps.kind == 1) {
          // ---- last Linear (both CTAs hold row 0 for all 2 NP points): sdf and the tanh slope ----
          wait_acc(0);
          wait_acc(1);
          if (q == 0) {
            const float bias0 = __ldg(Ps.bias);
#pragma unroll
            for (int g = 0; g < G; ++g) {
              uint32_t vm[GW], vc[GW];
              tmem_ld16(lane_base + ph * COL_PH + g * GW, vm);
              tmem_ld16(lane_base + ph * COL_PH + COL_CROSS + g * GW, vc);
              tmem_ld_wait();
              if (lane == 0) {
#pragma unroll
                for (int qq = 0; qq < GW; ++qq) {
                  const int n = ph * NP + g * GW + qq;
                  float y = (__uint_as_float(vm[qq]) + __uint_as_float(vc[qq])) * Ps.inv_scale + bias0, gg = 1.f;
                  if (use_tanh) { y = tanhf(y); gg *= 1.f - y * y; }
                  y = tanhf(y);
                  gg *= 1.f - y * y;
                  if (own && base + n < n_rows) sdf_out[base + n] = y;
                  gbuf[n] = gg;
                }
              }
              __syncwarp();
            }
          }
          tc_fence_before();
          asm volatile("bar.sync 1, 256;" ::: "memory");
          if (want_grad) {
            // backward of the last Linear is an outer product: delta[f][n] = W_last[f] * g[n] * mask[f][n]
            const int hidden = T.last_k;
            for (int j = 0; j < 2; ++j) {
              const int f = (2 * j + (int)rank) * 128 + t;
              const float w = f < hidden ? __ldg(T.last_w + f) * BWD_SCALE : 0.f;
              const unsigned long long mk = f < hidden ? mask_at(num_layers - 2, j) : 0ull;
              const uint32_t row = b_dst + (uint32_t)((f >> 3) * BCH + (f & 7) * 16);
#pragma unroll
              for (int g = 0; g < NP / 8; ++g) {
                float h[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) h[e] = ((mk >> (g * 8 + e)) & 1ull) ? w * gbuf[ph * NP + g * 8 + e] : 0.f;
                amax = fmaxf(amax, pack8_store_cluster(row, LO, g, h));
              }
              publish(j);
            }
          }
          continue;
        }
        if (Ps.kind == 3) {
          // ---- gradient with respect to the input rows (single block, both CTAs hold it): own points only ----
          wait_acc(0);
          wait_acc(1);
          if (q == 0 && own) {
            const int f = t;
#pragma unroll
            for (int g = 0; g < G; ++g) {
              uint32_t vm[GW], vc[GW];
              tmem_ld16(lane_base + ph * COL_PH + g * GW, vm);
              tmem_ld16(lane_base + ph * COL_PH + COL_CROSS + g * GW, vc);
              tmem_ld_wait();
              if (f < in0) {
#pragma unroll
                for (int qq = 0; qq < GW; ++qq)
                  dinp[f * NPP + ph * NP + g * GW + qq] += (__uint_as_float(vm[qq]) + __uint_as_float(vc[qq])) * Ps.inv_scale;
              }
            }
          }
          tc_fence_before();
          // the gradient of the concatenated input columns was accumulated by the CTA that owns those feature
          // lanes, for both point halves: add the peer's partial for this CTA's points
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster_release(peer_x);
          asm volatile("bar.sync 1, 256;" ::: "memory");
          mbar_wait_cluster(bar_x, x_phase);
          x_phase ^= 1;
          const uint32_t peer_dinp = mapa_u32(smem_u32(dinp), rank ^ 1u);
          for (int i = et; i < NP * in0; i += Q_NEPI) {
            const int n = i / in0, c = i - n * in0;
            const int np = (int)rank * NP + n;
            const float v = dinp[c * NPP + np] + ld_cluster_f32(peer_dinp + (uint32_t)((c * NPP + np) * 4));
            if (base + np < n_rows) dinput_out[(base + np) * in0 + c] = v;
          }
          continue;
        }
        const bool fwd = Ps.kind == 0;
        const int split = fwd ? Ps.rows : Ps.prev_rows;       // rows below: regular outputs
        const int nblocks = Ps.m_blocks >> 1;
        if (!kPipe) {                                  // in place: the operand may only change once nothing reads it any more
          wait_acc(0);
          wait_acc(1);
        }
        for (int j = 0; j < 2; ++j) {
          if (kPipe) wait_acc(j);
          if (j < nblocks) {
            const int f = (2 * j + (int)rank) * 128 + t;
            const int cls = f < split ? 0 : (f < split + Ps.cat_dim ? 1 : 2);
            const uint32_t tb = lane_base + (uint32_t)(j * 4 * NP + ph * COL_PH);
            const uint32_t row = b_dst + (uint32_t)((f >> 3) * BCH + (f & 7) * 16);
            const float bias = (fwd && cls == 0) ? __ldg(Ps.bias + f) : 0.f;
            const unsigned long long pmask = (!fwd && cls == 0) ? mask_at(Ps.layer - 1, j) : 0ull;
            const int cat_row = (Ps.cat_off + f - split) * NPP + ph * NP;
            unsigned long long mk = 0ull;
#pragma unroll
            for (int g = 0; g < G; ++g) {
              uint32_t vm[GW], vc[GW];
              tmem_ld16(tb + g * GW, vm);
              tmem_ld16(tb + COL_CROSS + g * GW, vc);
              tmem_ld_wait();
              float x[GW];
#pragma unroll
              for (int qq = 0; qq < GW; ++qq)
                x[qq] = (__uint_as_float(vm[qq]) + __uint_as_float(vc[qq])) * Ps.inv_scale;
              float h[GW];
              if (fwd) {
                if (cls == 0) {
#pragma unroll
                  for (int qq = 0; qq < GW; ++qq) {
                    const float y = x[qq] + bias;
                    const bool on = y > 0.f;
                    if (on) mk |= 1ull << (g * GW + qq);
                    h[qq] = on ? y * Ps.out_scale : 0.f;
                  }
                } else if (cls == 1) {                               // cat[x, input] feeds the next Linear
#pragma unroll
                  for (int qq = 0; qq < GW; ++qq) h[qq] = inp[cat_row + g * GW + qq] * Ps.out_scale;
                } else {
#pragma unroll
                  for (int qq = 0; qq < GW; ++qq) h[qq] = 0.f;
                }
              } else {
                if (cls == 0) {
#pragma unroll
                  for (int qq = 0; qq < GW; ++qq) h[qq] = ((pmask >> (g * GW + qq)) & 1ull) ? x[qq] * Ps.out_scale : 0.f;
                } else {
                  if (cls == 1) {                                    // gradient of the concatenated input columns
#pragma unroll
                    for (int qq = 0; qq < GW; ++qq) dinp[cat_row + g * GW + qq] += x[qq];
                  }
#pragma unroll
                  for (int qq = 0; qq < GW; ++qq) h[qq] = 0.f;
                }
              }
#pragma unroll
              for (int pk = 0; pk < PG; ++pk) {
                float h8[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) h8[e] = h[pk * 8 + e];
                amax = fmaxf(amax, pack8_store_cluster(row, LO, g * PG + pk, h8));
              }
            }
            if (fwd && want_grad) mask_at(Ps.layer, j) = mk;   // only the backward reads them
          }
          publish(j);
        }
      }
      // the next pair tile re-stages inp: every local epilogue thread must be past its last read
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }
    if (amax > 60000.f) atomicOr(overflow_flag, 1);   // a scaled operand left the fp16 range
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                 // no CTA may exit (or free TMEM) while its peer can still reach it
  if (warp == 9) tmem_dealloc_pair(tmem_base, TCOLS);
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// host: pass table + packed fp16 hi/lo weight tiles
// ---------------------------------------------------------------------------------------------
struct TcHostState {
  TcTable table;
  TcTable* table_dev;
  unsigned char* tiles_dev;
  int* overflow_dev;
  unsigned long long* mask_dev;
  size_t smem_bytes;
};

static float pow2_scale_for(const float* w, size_t n) {
  float mx = 0.f;
  for (size_t i = 0; i < n; ++i) mx = std::max(mx, std::fabs(w[i]));
  if (!(mx > 0.f) || !std::isfinite(mx)) return 1.f;
  int e;
  std::frexp(mx, &e);              // mx = m * 2^e, m in [0.5, 1)
  return std::ldexp(1.f, 7 - e);   // scaled max in [64, 128)
}

static void pack_tile(std::vector<__half>& out, size_t tile_index, const std::vector<float>& A, int rows, int K,
                      int lda, int mb, int kc, float scale) {
  __half* hi = out.data() + tile_index * (TILE_BYTES / 2);
  __half* lo = hi + TILE_HALF_BYTES / 2;
  for (int kc8 = 0; kc8 < KC / 8; ++kc8)
    for (int rg = 0; rg < 16; ++rg)
      for (int rr = 0; rr < 8; ++rr)
        for (int kk = 0; kk < 8; ++kk) {
          const int row = mb * 128 + rg * 8 + rr, k = kc * KC + kc8 * 8 + kk;
          const float v = (row < rows && k < K) ? A[(size_t)row * lda + k] * scale : 0.f;
          const __half h = __float2half_rn(v);
          const __half l = __float2half_rn(v - __half2float(h));
          const size_t o = ((size_t)(kc8 * 16 + rg) * 8 + rr) * 8 + kk;
          hi[o] = h;
          lo[o] = l;
        }
}

int build_tc_tables(sdfr_decoder* dec, const sdfr_decoder_spec* spec, const float* const* weights_host) {
  dec->tc.ok = 0;
  dec->tc_ptr = nullptr;
  const int NL = spec->num_layers, in0 = spec->latent_size + 3;
  if (NL > MAX_TC_LAYERS || in0 > KC) return SDFR_OK;
  if (spec->concat[NL - 1]) return SDFR_OK;   // a concatenating last Linear stays on the FFMA kernel
  for (int l = 0; l < NL; ++l) {
    if (spec->layer_norm[l]) return SDFR_OK;
    if (spec->in_dims[l] > 512 || spec->out_dims[l] > 512) return SDFR_OK;
  }
  int major = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dec->device);
  if (major != 10) return SDFR_OK;   // tcgen05 exists on sm_100 only
  const SmemPlan plan = make_plan<64>(NL, in0);
  if (plan.total > 227 * 1024) return SDFR_OK;

  TcHostState* st = new TcHostState();
  memset(&st->table, 0, sizeof(st->table));
  TcTable& T = st->table;
  T.num_layers = NL; T.in0 = in0; T.latent = spec->latent_size; T.use_tanh = spec->use_tanh;
  std::vector<float> wscale(NL);
  for (int l = 0; l < NL; ++l) wscale[l] = pow2_scale_for(weights_host[l], (size_t)spec->in_dims[l] * spec->out_dims[l]);

  struct HostPass { int kind, layer, rows, K; std::vector<float> A; int lda; float scale; };
  std::vector<HostPass> hp;
  // forward passes
  for (int l = 0; l < NL; ++l) {
    HostPass p;
    p.kind = l == NL - 1 ? 1 : 0; p.layer = l; p.rows = spec->out_dims[l]; p.K = spec->in_dims[l];
    p.A.assign(weights_host[l], weights_host[l] + (size_t)p.rows * p.K);
    p.lda = p.K; p.scale = wscale[l];
    hp.push_back(std::move(p));
  }
  // backward passes: W^T of layers NL-2 .. 0 (the last Linear's backward is an outer product in the epilogue)
  for (int l = NL - 2; l >= 0; --l) {
    HostPass p;
    p.kind = l == 0 ? 3 : 2; p.layer = l; p.rows = spec->in_dims[l]; p.K = spec->out_dims[l];
    p.A.resize((size_t)p.rows * p.K);
    for (int o = 0; o < spec->out_dims[l]; ++o)
      for (int i = 0; i < spec->in_dims[l]; ++i) p.A[(size_t)i * p.K + o] = weights_host[l][(size_t)o * spec->in_dims[l] + i];
    p.lda = p.K; p.scale = wscale[l];
    hp.push_back(std::move(p));
  }
  T.num_passes = (int)hp.size();
  std::vector<const float*> bias_dev(NL, nullptr);
  for (int l = 0; l < NL; ++l) bias_dev[l] = dec->dev.layer[l].bias;
  long long ntiles = 0;
  for (int pi = 0; pi < T.num_passes; ++pi) {
    HostPass& p = hp[pi];
    TcPassDev& D = T.pass[pi];
    D.kind = p.kind; D.layer = p.layer; D.rows = p.rows;
    D.m_blocks = (p.rows + 127) / 128; D.k_chunks = (p.K + KC - 1) / KC;
    D.bias = (p.kind <= 1) ? bias_dev[p.layer] : nullptr;
    D.cat_dim = 0; D.cat_off = 0; D.prev_rows = 0;
    const float in_scale = p.kind <= 1 ? ACT_SCALE : BWD_SCALE;
    D.inv_scale = 1.f / (p.scale * in_scale);
    D.out_scale = p.kind == 0 ? ACT_SCALE : BWD_SCALE;
    if (p.kind == 0 && p.layer + 1 < NL && spec->concat[p.layer + 1]) {      // the next Linear concatenates
      D.cat_dim = spec->concat[p.layer + 1] == 1 ? in0 : 3;
      D.cat_off = spec->concat[p.layer + 1] == 1 ? 0 : spec->latent_size;
      D.m_blocks = (p.rows + D.cat_dim + 127) / 128;                          // the concat rows are written by this pass
    }
    if (p.kind == 2) {
      D.prev_rows = spec->out_dims[p.layer - 1];
      if (spec->concat[p.layer]) {
        D.cat_dim = spec->concat[p.layer] == 1 ? in0 : 3;
        D.cat_off = spec->concat[p.layer] == 1 ? 0 : spec->latent_size;
      }
    }
    D.tile0 = ntiles;
    ntiles += (long long)D.m_blocks * D.k_chunks;
  }
  std::vector<__half> packed((size_t)ntiles * (TILE_BYTES / 2));
  long long tile = 0;
  int rc = SDFR_OK;
  for (int pi = 0; pi < T.num_passes; ++pi) {
    HostPass& p = hp[pi];
    const TcPassDev& D = T.pass[pi];
    for (int mb = 0; mb < D.m_blocks; ++mb)
      for (int kc = 0; kc < D.k_chunks; ++kc) pack_tile(packed, (size_t)tile++, p.A, p.rows, p.K, p.lda, mb, kc, p.scale);
  }
  T.tiles_per_point_tile = ntiles;
  T.last_k = spec->in_dims[NL - 1];
  T.last_w = dec->dev.layer[NL - 1].w;     // [out_pad=4][in_pad]: row 0 is the weight vector
  void* p = nullptr;
#define TC_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { set_error("%s: %s", #x, cudaGetErrorString(e_)); rc = SDFR_E_CUDA; } } while (0)
  TC_CUDA(cudaMalloc(&p, packed.size() * sizeof(__half)));
  if (rc == SDFR_OK) { dec->allocs.push_back(p); st->tiles_dev = reinterpret_cast<unsigned char*>(p); }
  if (rc == SDFR_OK) TC_CUDA(cudaMemcpy(p, packed.data(), packed.size() * sizeof(__half), cudaMemcpyHostToDevice));
  if (rc == SDFR_OK) { TC_CUDA(cudaMalloc(&p, sizeof(TcTable))); if (rc == SDFR_OK) { dec->allocs.push_back(p); st->table_dev = reinterpret_cast<TcTable*>(p); } }
  if (rc == SDFR_OK) TC_CUDA(cudaMemcpy(st->table_dev, &T, sizeof(TcTable), cudaMemcpyHostToDevice));
  if (rc == SDFR_OK) { TC_CUDA(cudaMalloc(&p, 64)); if (rc == SDFR_OK) { dec->allocs.push_back(p); st->overflow_dev = reinterpret_cast<int*>(p); TC_CUDA(cudaMemset(p, 0, 64)); } }
  if (rc == SDFR_OK) {
    const size_t mask_bytes = (size_t)(dec->sm_count > 0 ? dec->sm_count : 148) * (size_t)(NL - 1) * 512 * 8;
    TC_CUDA(cudaMalloc(&p, mask_bytes));
    if (rc == SDFR_OK) { dec->allocs.push_back(p); st->mask_dev = reinterpret_cast<unsigned long long*>(p); }
  }
  if (rc == SDFR_OK) TC_CUDA(cudaFuncSetAttribute(mlp_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.total));
  if (rc == SDFR_OK) TC_CUDA(cudaFuncSetAttribute(mlp_tc_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                  (int)make_plan<16>(NL, in0).total));
#undef TC_CUDA
  if (rc != SDFR_OK) { delete st; return rc; }
  st->smem_bytes = plan.total;
  dec->tc.ok = 1;
  dec->tc.num_passes = T.num_passes;
  dec->tc.num_tiles = ntiles;
  dec->tc.tiles = reinterpret_cast<const uint4*>(st->tiles_dev);
  dec->tc_ptr = reinterpret_cast<DecoderTc*>(st);   // opaque host state (freed with the decoder)
  return SDFR_OK;
}

#ifdef SDFR_TC_PROFILE
extern "C" int sdfr_debug_tc_prof(unsigned long long* out16, int reset) {
  if (out16) cudaMemcpyFromSymbol(out16, g_tc_prof, sizeof(unsigned long long) * 16);
  if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(g_tc_prof, z, sizeof(z)); }
  return 0;
}
#endif

void free_tc_tables(sdfr_decoder* dec) {
  if (dec->tc_ptr) delete reinterpret_cast<TcHostState*>(dec->tc_ptr);   // device memory is owned by dec->allocs
  dec->tc_ptr = nullptr;
}

// 1 when a scaled operand of the split-fp16 kernel has left the fp16 range since the last query
// (results of that launch are then invalid); synchronises the device.
int tc_overflow_flag(const sdfr_decoder* dec, int* flag) {
  *flag = 0;
  if (!dec->tc.ok || !dec->tc_ptr) return SDFR_OK;
  const TcHostState* st = reinterpret_cast<const TcHostState*>(dec->tc_ptr);
  SDFR_CUDA(cudaMemcpy(flag, st->overflow_dev, sizeof(int), cudaMemcpyDeviceToHost));
  if (*flag) SDFR_CUDA(cudaMemset(st->overflow_dev, 0, sizeof(int)));
  return SDFR_OK;
}

// Stream-ordered variant for callers that synchronise the stream anyway: enqueues the copy of the flag
// into *flag_host (0 when the decoder has no tensor-core tables); tc_overflow_reset clears it after a hit.
int tc_overflow_flag_enqueue(const sdfr_decoder* dec, int* flag_host, cudaStream_t s) {
  *flag_host = 0;
  if (!dec->tc.ok || !dec->tc_ptr) return SDFR_OK;
  const TcHostState* st = reinterpret_cast<const TcHostState*>(dec->tc_ptr);
  SDFR_CUDA(cudaMemcpyAsync(flag_host, st->overflow_dev, sizeof(int), cudaMemcpyDeviceToHost, s));
  return SDFR_OK;
}
const int* tc_overflow_ptr(const sdfr_decoder* dec) {
  if (!dec->tc.ok || !dec->tc_ptr) return nullptr;
  return reinterpret_cast<const TcHostState*>(dec->tc_ptr)->overflow_dev;
}
int tc_overflow_reset(const sdfr_decoder* dec, cudaStream_t s) {
  if (!dec->tc.ok || !dec->tc_ptr) return SDFR_OK;
  const TcHostState* st = reinterpret_cast<const TcHostState*>(dec->tc_ptr);
  SDFR_CUDA(cudaMemsetAsync(st->overflow_dev, 0, sizeof(int), s));
  return SDFR_OK;
}

// CTA pairs the band-pair kernel can keep resident (0: the kernel cannot run here), queried once per tile width and
// input width (the shared-memory plan grows with the decoder's input).
template <int NP>
static int band_pair_slots(int in0) {
  static int cache[5] = {-1, -1, -1, -1, -1};
  static int smem_max = -1;
  if (in0 <= 0 || in0 > KC) return 0;
  const int key = (in0 + 7) / 8;
  if (cache[key] < 0) {
    cache[key] = 0;
    if (smem_max < 0) {
      int devid = 0;
      smem_max = 0;
      if (cudaGetDevice(&devid) != cudaSuccess ||
          cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, devid) != cudaSuccess ||
          cudaFuncSetAttribute(mlp_tc_band_pair_kernel<NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max) != cudaSuccess)
        smem_max = 0;
    }
    const BandPairPlan plan = make_band_pair_plan<NP>(in0);
    if ((int)plan.total <= smem_max) {
      cudaLaunchConfig_t q;
      memset(&q, 0, sizeof(q));
      q.gridDim = dim3((unsigned)(2 * 64));
      q.blockDim = dim3(Q_THREADS);
      q.dynamicSmemBytes = plan.total;
      cudaLaunchAttribute qa[1];
      qa[0].id = cudaLaunchAttributeClusterDimension;
      qa[0].val.clusterDim.x = 2; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
      q.attrs = qa; q.numAttrs = 1;
      int nc = 0;
      if (cudaOccupancyMaxActiveClusters(&nc, mlp_tc_band_pair_kernel<NP>, &q) == cudaSuccess && nc >= 1) cache[key] = nc;
    }
    cudaGetLastError();
  }
  return cache[key];
}

// The band-pair kernel takes pass tables whose hidden passes are an even number of 128-feature blocks and whose
// single-row passes (last Linear, gradient of the first) are one block; everything else stays on mlp_tc_kernel<16>.
static bool band_pair_ok(const sdfr_decoder* dec, const TcHostState* st, bool want_grad) {
  static int enabled = -1;
  if (enabled < 0) { const char* e = getenv("SDFR_BAND_PAIR"); enabled = e ? atoi(e) : 1; }   // 0: A/B runs of the single-CTA kernel
  if (!enabled) return false;
  const int npass = want_grad ? st->table.num_passes : st->table.num_layers;
  for (int p = 0; p < npass; ++p) {
    const TcPassDev& ps = st->table.pass[p];
    const bool single = ps.kind == 1 || ps.kind == 3;
    if (single ? ps.m_blocks != 1 : (ps.m_blocks != 2 && ps.m_blocks != 4)) return false;
  }
  return true;
}

template <int NP>
static int launch_band_pair(const sdfr_decoder* dec, const TcHostState* st, const MlpInputs& in, float* sdf, float* dinput,
                            unsigned long long* masks, cudaStream_t s) {
  const long long pair_tiles = (in.n + 2 * NP - 1) / (2 * NP);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(2 * std::min<long long>(pair_tiles, band_pair_slots<NP>(dec->dev.in0))));
  cfg.blockDim = dim3(Q_THREADS);
  cfg.dynamicSmemBytes = make_band_pair_plan<NP>(dec->dev.in0).total;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  SDFR_CUDA(cudaLaunchKernelEx(&cfg, mlp_tc_band_pair_kernel<NP>, (const TcTable*)st->table_dev,
                               (const unsigned char*)st->tiles_dev, in, sdf, dinput, st->overflow_dev, masks));
  SDFR_LAUNCH_CHECK();
  return SDFR_OK;
}

// Forward + input gradient (or forward only when dinput is null) at full fp32-equivalent precision.
int launch_mlp_tc(const sdfr_decoder* dec, const MlpInputs& in, float* sdf, float* dinput, cudaStream_t s) {
  SDFR_REQUIRE(dec->tc.ok && dec->tc_ptr, SDFR_E_UNSUPPORTED,
               "tcgen05 MLP kernel does not cover this decoder (needs sm_100, widths <= 512, no LayerNorm, "
               "latent+3 <= 32, <= 9 layers)");
  if (in.n <= 0) return SDFR_OK;
  const TcHostState* st = reinterpret_cast<const TcHostState*>(dec->tc_ptr);
  // 16-point tiles when the caller expects a short row list (MlpInputs::small_tiles: the band pass of a
  // few detections), so that ~2 000 rows still fill the machine; 64-point tiles otherwise.
  const int sms = dec->sm_count > 0 ? dec->sm_count : 148;
  const bool small = in.small_tiles || (!in.count_dev && in.n <= 16LL * sms);
  const int np = small ? 16 : 64;
  const long long point_tiles = (in.n + np - 1) / np;
  const int grid = (int)std::min<long long>(point_tiles, sms);
  // Weight stages per issuer iteration.  16-point tiles: 4 (210 -> 173 us for ~1 850 rows).  64-point tiles: 2 when
  // a CTA streams several tiles (1.68 -> 1.58 ms for the 64 000-point forward+gradient sweep); a CTA with a single
  // 64-point tile waits on the weight stream and loses 10 % by waiting for two stages of its 5-stage ring.
  const int group = np == 16 ? 4 : (point_tiles > grid ? 2 : 1);
  unsigned long long* masks = in.mask_scratch ? in.mask_scratch : st->mask_dev;
  if (band_pair_ok(dec, st, dinput != nullptr)) {
    // The row count lives on the device and may be anything (trace mode's Newton rounds): both tilings are enqueued,
    // each with a window on the count, and the one whose window it misses returns at once (~3 us)
    if (in.adaptive_tiles && in.count_dev && band_pair_slots<16>(dec->dev.in0) > 0 && band_pair_slots<64>(dec->dev.in0) > 0) {
      MlpInputs lo = in, hi = in;
      lo.count_hi = hi.count_lo = 2 * 32 * band_pair_slots<16>(dec->dev.in0);      // two rounds of 2 x 16 ~ one of 2 x 64
      int rc = launch_band_pair<16>(dec, st, lo, sdf, dinput, masks, s);
      return rc ? rc : launch_band_pair<64>(dec, st, hi, sdf, dinput, masks, s);
    }
    // CTA pairs, M = 256 features per instruction: 2 x 16 points per pair for short row lists, 2 x 64 otherwise
    if (small && band_pair_slots<16>(dec->dev.in0) > 0) return launch_band_pair<16>(dec, st, in, sdf, dinput, masks, s);
    if (!small && band_pair_slots<64>(dec->dev.in0) > 0) return launch_band_pair<64>(dec, st, in, sdf, dinput, masks, s);
  }
  if (small)
    mlp_tc_kernel<16><<<grid, NTHREADS, make_plan<16>(dec->dev.num_layers, dec->dev.in0).total, s>>>(
        st->table_dev, st->tiles_dev, in, sdf, dinput, st->overflow_dev, masks, point_tiles, group);
  else
    mlp_tc_kernel<64><<<grid, NTHREADS, st->smem_bytes, s>>>(st->table_dev, st->tiles_dev, in, sdf, dinput,
                                                             st->overflow_dev, masks, point_tiles, group);
  SDFR_LAUNCH_CHECK();
  return SDFR_OK;
}

size_t mlp_tc_mask_scratch_bytes(const sdfr_decoder* dec) {
  if (!dec->tc.ok) return 0;
  return (size_t)(dec->sm_count > 0 ? dec->sm_count : 148) * (size_t)(dec->dev.num_layers - 1) * 512 * 8;
}

// Forward only at fp16 operand precision (hi halves): the lattice pass of the fused engine.  CTA-pair kernel for
// stock-like pass tables (every hidden pass an even number of 128-feature blocks, last Linear a single row);
// the wide-tile kernel for every other table the tensor-core path accepts.
static int g_pair_slots = -1, g_pair_smem_max = 0;

// CTA pairs the lattice-pass kernel can keep resident, queried once (0: it cannot run here)
static int coarse_pair_slots() {
  if (g_pair_slots < 0) {
    g_pair_slots = 0;
    // opt in to the device maximum once: the plan grows with the decoder's input width
    int devid = 0;
    PairPlan plan = make_pair_plan(3 + 3);              // the stock input width, for the occupancy query
    if (cudaGetDevice(&devid) == cudaSuccess &&
        cudaDeviceGetAttribute(&g_pair_smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, devid) == cudaSuccess &&
        (int)plan.total <= g_pair_smem_max &&
        cudaFuncSetAttribute(mlp_tc_coarse_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, g_pair_smem_max) == cudaSuccess) {
      cudaLaunchConfig_t q;
      memset(&q, 0, sizeof(q));
      q.gridDim = dim3((unsigned)(2 * 64));
      q.blockDim = dim3(P_THREADS);
      q.dynamicSmemBytes = plan.total;
      cudaLaunchAttribute qa[1];
      qa[0].id = cudaLaunchAttributeClusterDimension;
      qa[0].val.clusterDim.x = 2; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
      q.attrs = qa; q.numAttrs = 1;
      int nc = 0;
      if (cudaOccupancyMaxActiveClusters(&nc, mlp_tc_coarse_pair_kernel, &q) == cudaSuccess && nc >= 1) g_pair_slots = nc;
    }
    cudaGetLastError();
  }
  return g_pair_slots;
}

// per decoder: does its pass table fit the pair kernel (every hidden pass an even number of 128-feature blocks,
// last Linear a single row)?
static bool coarse_pair_ok(const sdfr_decoder* dec) {
  if (!dec->tc.ok || !dec->tc_ptr) return false;
  const TcHostState* st = reinterpret_cast<const TcHostState*>(dec->tc_ptr);
  bool ok = coarse_pair_slots() > 0 && (int)make_pair_plan(dec->dev.in0).total <= g_pair_smem_max;
  for (int p = 0; ok && p < st->table.num_layers; ++p) {
    const TcPassDev& ps = st->table.pass[p];
    const bool last = ps.kind == 1;
    if (last ? ps.m_blocks != 1 : (ps.m_blocks != 2 && ps.m_blocks != 4)) ok = false;
    if (ps.k_chunks != 1 && (ps.k_chunks & 1)) ok = false;
  }
  return ok;
}

// march mode of trace.cu lives in the pair kernel only
bool mlp_tc_march_ok(const sdfr_decoder* dec) { return coarse_pair_ok(dec); }
// rows one round of the lattice-pass grid evaluates (every resident CTA pair one tile)
int mlp_tc_round_rows(const sdfr_decoder* dec) { return coarse_pair_ok(dec) ? coarse_pair_slots() * 2 * P_PTS : 0; }

int launch_mlp_tc_coarse(const sdfr_decoder* dec, const MlpInputs& in, float* sdf, cudaStream_t s) {
  SDFR_REQUIRE(dec->tc.ok && dec->tc_ptr, SDFR_E_UNSUPPORTED, "tcgen05 MLP kernel does not cover this decoder");
  if (in.n <= 0) return SDFR_OK;
  const TcHostState* st = reinterpret_cast<const TcHostState*>(dec->tc_ptr);
  const int sms = dec->sm_count > 0 ? dec->sm_count : 148;
  const bool pair_ok = coarse_pair_ok(dec);
  const int pair_slots = coarse_pair_slots();
  SDFR_REQUIRE(pair_ok || !in.march, SDFR_E_UNSUPPORTED, "march mode needs the CTA-pair lattice kernel");
  if (pair_ok) {
    const long long pair_tiles = (in.n + 2 * P_PTS - 1) / (2 * P_PTS);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(2 * std::min<long long>(pair_tiles, pair_slots)));
    cfg.blockDim = dim3(P_THREADS);
    cfg.dynamicSmemBytes = make_pair_plan(dec->dev.in0).total;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    SDFR_CUDA(cudaLaunchKernelEx(&cfg, mlp_tc_coarse_pair_kernel, (const TcTable*)st->table_dev,
                                 (const unsigned char*)st->tiles_dev, in, sdf));
    SDFR_LAUNCH_CHECK();
    return SDFR_OK;
  }
  // wide tiles: as wide as the CTA can hold next to the weight ring (112 points with 32 KB stages), shrunk so
  // that all rounds of tiles are equally full
  const int max_npt = 112;
  const long long rounds = ((in.n + max_npt - 1) / max_npt + sms - 1) / sms;
  const long long w = (in.n + rounds * sms - 1) / (rounds * sms);
  const int npt = (int)std::min<long long>(max_npt, ((w + 15) / 16) * 16);
  const long long tiles_n = (in.n + npt - 1) / npt;
  const int grid = (int)std::min<long long>(tiles_n, sms);
  static bool wide_attr_set = false;
  if (!wide_attr_set) {
    SDFR_CUDA(cudaFuncSetAttribute(mlp_tc_coarse_wide_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)make_wide_plan<4>(KC, 112).total));
    wide_attr_set = true;
  }
  mlp_tc_coarse_wide_kernel<4><<<grid, W_THREADS, make_wide_plan<4>(dec->dev.in0, npt).total, s>>>(
      st->table_dev, st->tiles_dev, in, sdf, npt);
  SDFR_LAUNCH_CHECK();
  return SDFR_OK;
}

}  // namespace sdfr
