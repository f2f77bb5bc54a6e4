// DeepSDF decoder, fp32 CUDA-core kernel: sdf = f(latent, x) and the exact
// gradient d sdf / d [latent, x] for a tile of points per CTA.
//
// replaces: Decoder.forward (sdfrenderer/deepsdf/networks/deep_sdf_decoder_scale.py:78-114)
//           and the autograd pass pred_sdf_grid.sum().backward() that yields the
//           normals (sdfrenderer/grid.py:55-56), restricted to the input gradient
//           (the reference also computes and discards dW for every layer).
//
// One CTA owns TP = 32 points; a warp owns 4 complete activation rows, so
// LayerNorm statistics are warp shuffles.  The forward keeps only ReLU sign
// bits (one u64 per thread per layer); the backward walks W instead of W^T.
// Because the backward yields d sdf / d latent for every point too, the refine
// loop needs no third MLP pass ("MLP backward #2").
//
// This kernel is the general-spec path (any dims <= 512, LayerNorm, xyz_in_all,
// use_tanh); the tensor-core kernel in mlp_tc.cu covers the stock spec class.
#include "common.cuh"

namespace sdfr {

namespace {

constexpr int TP = 32;        // points per CTA
constexpr int NT = 256;       // threads
constexpr int PPW = 4;        // points per warp
constexpr int NG = 4;         // float4 column groups per lane (4*32*4 = 512 columns)

struct SmemLayout {
  int act_stride;     // floats per activation row
  int in_stride;      // floats per input row
  size_t off_act0, off_act1, off_inp, off_dinp, off_mask, off_rstd, off_g, total;
};

__host__ __device__ inline SmemLayout make_layout(int max_width, int in0, int num_layers) {
  SmemLayout s;
  s.act_stride = max_width;
  s.in_stride = (in0 + 3) & ~3;
  size_t o = 0;
  s.off_act0 = o; o += (size_t)TP * s.act_stride * 4;
  s.off_act1 = o; o += (size_t)TP * s.act_stride * 4;
  s.off_inp = o;  o += (size_t)TP * s.in_stride * 4;
  s.off_dinp = o; o += (size_t)TP * s.in_stride * 4;
  s.off_mask = o; o += (size_t)num_layers * NT * 8;
  s.off_rstd = o; o += (size_t)num_layers * TP * 4;
  s.off_g = o;    o += (size_t)TP * 4;
  s.total = o;
  return s;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// acc[pp][j*4+e] += sum_k act[(row0+pp)][k] * B[k][ (lane+32j)*4 + e ]
__device__ __forceinline__ void gemm_rows(const float* __restrict__ act, int act_stride, int row0, int kpad,
                                          const float* __restrict__ B, int ldb, int npad, int lane,
                                          float (&acc)[PPW][NG * 4]) {
  bool valid[NG];
#pragma unroll
  for (int j = 0; j < NG; ++j) valid[j] = (lane + 32 * j) * 4 < npad;
#pragma unroll
  for (int pp = 0; pp < PPW; ++pp)
#pragma unroll
    for (int c = 0; c < NG * 4; ++c) acc[pp][c] = 0.f;

  const float* arow = act + (size_t)row0 * act_stride;
#pragma unroll 2
  for (int k0 = 0; k0 < kpad; k0 += 4) {
    float4 a[PPW];
#pragma unroll
    for (int pp = 0; pp < PPW; ++pp) a[pp] = *reinterpret_cast<const float4*>(arow + pp * act_stride + k0);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      float4 b[NG];
      const float* brow = B + (size_t)(k0 + kk) * ldb + lane * 4;
#pragma unroll
      for (int j = 0; j < NG; ++j)
        b[j] = valid[j] ? __ldg(reinterpret_cast<const float4*>(brow + 128 * j)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int pp = 0; pp < PPW; ++pp) {
        const float av = kk == 0 ? a[pp].x : kk == 1 ? a[pp].y : kk == 2 ? a[pp].z : a[pp].w;
#pragma unroll
        for (int j = 0; j < NG; ++j) {
          acc[pp][j * 4 + 0] = fmaf(av, b[j].x, acc[pp][j * 4 + 0]);
          acc[pp][j * 4 + 1] = fmaf(av, b[j].y, acc[pp][j * 4 + 1]);
          acc[pp][j * 4 + 2] = fmaf(av, b[j].z, acc[pp][j * 4 + 2]);
          acc[pp][j * 4 + 3] = fmaf(av, b[j].w, acc[pp][j * 4 + 3]);
        }
      }
    }
  }
}

// Narrow outputs (npad <= 16): lanes stride over k, shuffle-reduce at the end.
// out[pp][c] valid on every lane afterwards.
__device__ __forceinline__ void gemm_rows_narrow(const float* __restrict__ act, int act_stride, int row0, int kpad,
                                                 const float* __restrict__ B, int ldb, int npad, int lane,
                                                 float (&out)[PPW][16]) {
#pragma unroll
  for (int pp = 0; pp < PPW; ++pp)
#pragma unroll
    for (int c = 0; c < 16; ++c) out[pp][c] = 0.f;
  const int ng = npad >> 2;
  for (int k = lane; k < kpad; k += 32) {
    float a[PPW];
#pragma unroll
    for (int pp = 0; pp < PPW; ++pp) a[pp] = act[(size_t)(row0 + pp) * act_stride + k];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (j < ng) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(B + (size_t)k * ldb + j * 4));
#pragma unroll
        for (int pp = 0; pp < PPW; ++pp) {
          out[pp][j * 4 + 0] = fmaf(a[pp], b.x, out[pp][j * 4 + 0]);
          out[pp][j * 4 + 1] = fmaf(a[pp], b.y, out[pp][j * 4 + 1]);
          out[pp][j * 4 + 2] = fmaf(a[pp], b.z, out[pp][j * 4 + 2]);
          out[pp][j * 4 + 3] = fmaf(a[pp], b.w, out[pp][j * 4 + 3]);
        }
      }
    }
  }
#pragma unroll
  for (int pp = 0; pp < PPW; ++pp)
#pragma unroll
    for (int c = 0; c < 16; ++c) out[pp][c] = warp_sum(out[pp][c]);
}

__global__ void __launch_bounds__(NT, 1)
mlp_ffma_kernel(const DecoderDev* __restrict__ decp, MlpInputs in, float* __restrict__ sdf_out,
                float* __restrict__ dinput_out, float* __restrict__ ln_scratch, long long ln_scratch_per_cta) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const DecoderDev& dec = *decp;
  const SmemLayout L = make_layout(dec.max_width, dec.in0, dec.num_layers);
  float* act[2] = {reinterpret_cast<float*>(smem_raw + L.off_act0), reinterpret_cast<float*>(smem_raw + L.off_act1)};
  float* inp = reinterpret_cast<float*>(smem_raw + L.off_inp);
  float* dinp = reinterpret_cast<float*>(smem_raw + L.off_dinp);
  unsigned long long* maskbuf = reinterpret_cast<unsigned long long*>(smem_raw + L.off_mask);
  float* rstd_buf = reinterpret_cast<float*>(smem_raw + L.off_rstd);
  float* gbuf = reinterpret_cast<float*>(smem_raw + L.off_g);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int row0 = warp * PPW;
  const long long base = (long long)blockIdx.x * TP;
  const long long n_rows = mlp_rows(in);
  if (base >= n_rows) return;            // launched for the capacity; the row count lives on the device
  const int in0 = dec.in0, Ls = dec.latent_size, NL = dec.num_layers;
  const int AS = L.act_stride, IS = L.in_stride;
  float* xhat_scratch = ln_scratch ? ln_scratch + (size_t)blockIdx.x * ln_scratch_per_cta : nullptr;

  // ---- stage the inputs ---------------------------------------------------
  for (int i = tid; i < TP * IS; i += NT) {
    const int p = i / IS, c = i - p * IS;
    const long long gi = base + p;
    float v = 0.f;
    if (gi < n_rows && c < in0) {
      const long long src = in.index ? (long long)in.index[gi] : gi;
      if (in.inputs) {
        v = in.inputs[src * in0 + c];
      } else {
        const long long b = src / in.points_per_batch, k = src - b * in.points_per_batch;
        if (c < Ls) {
          v = in.latent_unit[b * Ls + c];
        } else {
          float x, y, z;
          lattice_point(in.lattice, k, x, y, z);
          v = (c - Ls) == 0 ? x : (c - Ls) == 1 ? y : z;
        }
      }
    }
    inp[i] = v;
    dinp[i] = 0.f;
  }
  __syncthreads();
  {
    const int kpad0 = dec.layer[0].in_pad;
    for (int i = tid; i < TP * kpad0; i += NT) {
      const int p = i / kpad0, c = i - p * kpad0;
      act[0][(size_t)p * AS + c] = c < in0 ? inp[p * IS + c] : 0.f;
    }
  }
  __syncthreads();

  int cur = 0;
  // ---- forward --------------------------------------------------------------
  for (int l = 0; l < NL; ++l) {
    const LayerDev& Ly = dec.layer[l];
    if (Ly.concat) {   // cat[x, input] / cat[x, xyz] before this Linear (decoder.py:90-93)
      const int prev = dec.layer[l - 1].out_dim;
      const int cdim = Ly.concat == 1 ? in0 : 3, coff = Ly.concat == 1 ? 0 : Ls;
      const int span = Ly.in_pad - prev;
      for (int i = tid; i < TP * span; i += NT) {
        const int p = i / span, c = i - p * span;
        act[cur][(size_t)p * AS + prev + c] = c < cdim ? inp[p * IS + coff + c] : 0.f;
      }
      __syncthreads();
    }
    if (l < NL - 1) {
      float acc[PPW][NG * 4];
      gemm_rows(act[cur], AS, row0, Ly.in_pad, Ly.wt, Ly.out_pad, Ly.out_pad, lane, acc);
      // bias
#pragma unroll
      for (int j = 0; j < NG; ++j) {
        const int n0 = (lane + 32 * j) * 4;
        if (n0 < Ly.out_pad) {
          const float4 bv = __ldg(reinterpret_cast<const float4*>(Ly.bias + n0));
#pragma unroll
          for (int pp = 0; pp < PPW; ++pp) {
            acc[pp][j * 4 + 0] += bv.x; acc[pp][j * 4 + 1] += bv.y;
            acc[pp][j * 4 + 2] += bv.z; acc[pp][j * 4 + 3] += bv.w;
          }
        }
      }
      if (Ly.layer_norm) {   // nn.LayerNorm(out) (decoder.py:99-101), biased variance, eps 1e-5
        const float inv_n = 1.f / (float)Ly.out_dim;
#pragma unroll
        for (int pp = 0; pp < PPW; ++pp) {
          float s = 0.f;
#pragma unroll
          for (int j = 0; j < NG; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if ((lane + 32 * j) * 4 + e < Ly.out_dim) s += acc[pp][j * 4 + e];
          const float mean = warp_sum(s) * inv_n;
          float q = 0.f;
#pragma unroll
          for (int j = 0; j < NG; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if ((lane + 32 * j) * 4 + e < Ly.out_dim) {
                const float d = acc[pp][j * 4 + e] - mean;
                q += d * d;
              }
          const float rstd = rsqrtf(warp_sum(q) * inv_n + 1e-5f);
          if (lane == 0) rstd_buf[l * TP + row0 + pp] = rstd;
#pragma unroll
          for (int j = 0; j < NG; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int n = (lane + 32 * j) * 4 + e;
              if (n < Ly.out_dim) {
                const float xh = (acc[pp][j * 4 + e] - mean) * rstd;
                if (xhat_scratch) xhat_scratch[((size_t)l * TP + row0 + pp) * kMaxWidthFFMA + n] = xh;
                acc[pp][j * 4 + e] = xh * __ldg(Ly.ln_w + n) + __ldg(Ly.ln_b + n);
              } else {
                acc[pp][j * 4 + e] = 0.f;
              }
            }
        }
      }
      // ReLU + sign bits + store
      unsigned long long bits = 0ull;
      float* dst = act[cur ^ 1];
#pragma unroll
      for (int pp = 0; pp < PPW; ++pp)
#pragma unroll
        for (int j = 0; j < NG; ++j) {
          const int n0 = (lane + 32 * j) * 4;
          float4 o;
          float* ov = reinterpret_cast<float*>(&o);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float v = acc[pp][j * 4 + e];
            const bool on = v > 0.f;
            if (on) bits |= 1ull << (pp * 16 + j * 4 + e);
            ov[e] = on ? v : 0.f;
          }
          if (n0 < Ly.out_pad) *reinterpret_cast<float4*>(dst + (size_t)(row0 + pp) * AS + n0) = o;
        }
      maskbuf[(size_t)l * NT + tid] = bits;
      __syncthreads();
      cur ^= 1;
    } else {
      // last Linear (narrow), optional tanh, final tanh (decoder.py:96-97,106-107)
      float y[PPW][16];
      gemm_rows_narrow(act[cur], AS, row0, Ly.in_pad, Ly.wt, Ly.out_pad, Ly.out_pad, lane, y);
      if (lane == 0) {
        const float b = __ldg(Ly.bias);
#pragma unroll
        for (int pp = 0; pp < PPW; ++pp) {
          float v = y[pp][0] + b, g = 1.f;
          if (dec.use_tanh) { v = tanhf(v); g *= 1.f - v * v; }
          v = tanhf(v);
          g *= 1.f - v * v;
          const long long gi = base + row0 + pp;
          if (gi < n_rows) sdf_out[gi] = v;
          gbuf[row0 + pp] = g;
        }
      }
      __syncthreads();
    }
  }
  if (!dinput_out) return;

  // ---- backward: gradient of sdf with respect to the input row ----------------
  {
    const LayerDev& Ly = dec.layer[NL - 1];
    for (int i = tid; i < TP * Ly.out_pad; i += NT) {
      const int p = i / Ly.out_pad, c = i - p * Ly.out_pad;
      act[cur][(size_t)p * AS + c] = c == 0 ? gbuf[p] : 0.f;
    }
    __syncthreads();
  }
  for (int l = NL - 1; l >= 0; --l) {
    const LayerDev& Ly = dec.layer[l];
    // act[cur][p][0..out_pad) = d sdf / d (Linear_l output)
    if (l == 0) {
      if (Ly.in_pad <= 16) {
        float d[PPW][16];
        gemm_rows_narrow(act[cur], AS, row0, Ly.out_pad, Ly.w, Ly.in_pad, Ly.in_pad, lane, d);
        if (lane == 0) {
#pragma unroll
          for (int pp = 0; pp < PPW; ++pp)
#pragma unroll
            for (int c = 0; c < 16; ++c)
              if (c < in0) dinp[(row0 + pp) * IS + c] += d[pp][c];
        }
      } else {
        float acc[PPW][NG * 4];
        gemm_rows(act[cur], AS, row0, Ly.out_pad, Ly.w, Ly.in_pad, Ly.in_pad, lane, acc);
#pragma unroll
        for (int pp = 0; pp < PPW; ++pp)
#pragma unroll
          for (int j = 0; j < NG; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int n = (lane + 32 * j) * 4 + e;
              if (n < in0) dinp[(row0 + pp) * IS + n] += acc[pp][j * 4 + e];
            }
      }
      break;   // layer 0 has no predecessor
    }
    float acc[PPW][NG * 4];
    gemm_rows(act[cur], AS, row0, Ly.out_pad, Ly.w, Ly.in_pad, Ly.in_pad, lane, acc);
    const LayerDev& Pv = dec.layer[l - 1];
    const int prev = Pv.out_dim;
    const int cdim = Ly.concat == 1 ? in0 : (Ly.concat == 2 ? 3 : 0), coff = Ly.concat == 1 ? 0 : Ls;
    const unsigned long long bits = maskbuf[(size_t)(l - 1) * NT + tid];
    // concat columns feed the input gradient; the rest go through ReLU (and LayerNorm) of layer l-1
#pragma unroll
    for (int pp = 0; pp < PPW; ++pp)
#pragma unroll
      for (int j = 0; j < NG; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int n = (lane + 32 * j) * 4 + e;
          float v = acc[pp][j * 4 + e];
          if (n >= prev) {
            if (n < prev + cdim) dinp[(row0 + pp) * IS + coff + (n - prev)] += v;
            v = 0.f;
          } else if (!((bits >> (pp * 16 + j * 4 + e)) & 1ull)) {
            v = 0.f;
          }
          acc[pp][j * 4 + e] = v;
        }
    if (Pv.layer_norm) {
      const float inv_n = 1.f / (float)prev;
#pragma unroll
      for (int pp = 0; pp < PPW; ++pp) {
        float s1 = 0.f, s2 = 0.f;
        float xh[NG * 4];
#pragma unroll
        for (int j = 0; j < NG; ++j)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int n = (lane + 32 * j) * 4 + e;
            float x = 0.f, gg = 0.f;
            if (n < prev) {
              x = xhat_scratch[((size_t)(l - 1) * TP + row0 + pp) * kMaxWidthFFMA + n];
              gg = acc[pp][j * 4 + e] * __ldg(Pv.ln_w + n);
            }
            xh[j * 4 + e] = x;
            acc[pp][j * 4 + e] = gg;
            s1 += gg;
            s2 += gg * x;
          }
        const float m1 = warp_sum(s1) * inv_n, m2 = warp_sum(s2) * inv_n;
        const float rstd = rstd_buf[(l - 1) * TP + row0 + pp];
#pragma unroll
        for (int j = 0; j < NG; ++j)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int n = (lane + 32 * j) * 4 + e;
            acc[pp][j * 4 + e] = n < prev ? rstd * (acc[pp][j * 4 + e] - m1 - xh[j * 4 + e] * m2) : 0.f;
          }
      }
    }
    float* dst = act[cur ^ 1];
#pragma unroll
    for (int pp = 0; pp < PPW; ++pp)
#pragma unroll
      for (int j = 0; j < NG; ++j) {
        const int n0 = (lane + 32 * j) * 4;
        if (n0 < Pv.out_pad)
          *reinterpret_cast<float4*>(dst + (size_t)(row0 + pp) * AS + n0) =
              make_float4(acc[pp][j * 4 + 0], acc[pp][j * 4 + 1], acc[pp][j * 4 + 2], acc[pp][j * 4 + 3]);
      }
    __syncthreads();
    cur ^= 1;
  }
  __syncthreads();
  for (int i = tid; i < TP * in0; i += NT) {
    const int p = i / in0, c = i - p * in0;
    const long long gi = base + p;
    if (gi < n_rows) dinput_out[gi * in0 + c] = dinp[p * IS + c];
  }
}

}  // namespace

int launch_mlp_ffma(const sdfr_decoder* dec, const MlpInputs& in, float* sdf, float* dinput, cudaStream_t s) {
  if (in.n <= 0) return SDFR_OK;
  const DecoderDev& d = dec->dev;
  SDFR_REQUIRE(d.max_width <= kMaxWidthFFMA, SDFR_E_UNSUPPORTED, "layer width %d > %d unsupported", d.max_width,
               kMaxWidthFFMA);
  const SmemLayout L = make_layout(d.max_width, d.in0, d.num_layers);
  SDFR_REQUIRE(L.total <= 227 * 1024, SDFR_E_UNSUPPORTED, "decoder needs %zu B of shared memory (> 227 KB)", L.total);
  const long long tiles = (in.n + TP - 1) / TP;
  bool any_ln = false;
  for (int l = 0; l < d.num_layers; ++l) any_ln |= d.layer[l].layer_norm != 0;
  float* scratch = nullptr;
  long long per_cta = 0;
  if (any_ln && dinput) {
    per_cta = (long long)d.num_layers * TP * kMaxWidthFFMA;
    const size_t need = (size_t)tiles * per_cta * sizeof(float);
    SDFR_REQUIRE(need <= ((size_t)8 << 30), SDFR_E_CAPACITY, "LayerNorm backward scratch of %zu B too large", need);
    sdfr_decoder* md = const_cast<sdfr_decoder*>(dec);
    if (md->scratch_bytes < need) {
      if (md->scratch) SDFR_CUDA(cudaFree(md->scratch));
      md->scratch = nullptr;
      md->scratch_bytes = 0;
      SDFR_CUDA(cudaMalloc(&md->scratch, need));
      md->scratch_bytes = need;
    }
    scratch = md->scratch;
  }
  static bool attr_set = false;
  if (!attr_set) {
    SDFR_CUDA(cudaFuncSetAttribute(mlp_ffma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  mlp_ffma_kernel<<<(unsigned)tiles, NT, L.total, s>>>(dec->dev_ptr, in, sdf, dinput, scratch, per_cta);
  SDFR_LAUNCH_CHECK();
  return SDFR_OK;
}

}  // namespace sdfr
