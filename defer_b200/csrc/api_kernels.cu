// api_kernels.cu - per-kernel C-ABI entry points (raw device pointers) for parity tests and ncu.
#include <stdlib.h>

#include <vector>

#include "common.cuh"
#include "conv_umma.cuh"

using namespace defer;

extern "C" {

int defer_k_conv(int fmt, int backend, const void* x, int x_is_f32, const float* w_hwio, const float* scale,
                 const float* shift, const void* residual, void* y, int n, int h, int w, int cin, int cout, int kh, int kw,
                 int sh, int sw, int pad_t, int pad_l, int pad_b, int pad_r, uint32_t flags, void* stream) {
  DEFER_CHECK(x && w_hwio && y, "k_conv: null pointer");
  DEFER_CHECK(fmt >= 0 && fmt <= 2, "k_conv: bad fmt");
  cudaStream_t st = (cudaStream_t)stream;
  int ho = (h + pad_t + pad_b - kh) / sh + 1, wo = (w + pad_l + pad_r - kw) / sw + 1;
  DEFER_CHECK(ho >= 1 && wo >= 1, "k_conv: empty output");
  if (residual) flags |= DEFER_FLAG_RESIDUAL;
  if (backend >= 4 && backend <= 7) {
    // streaming persistent kernel (conv_stream_kernel): 4 = 64-wide N tiles, 5 = 128-wide, 6 / 7 = the same with the
    // per-thread (peer-memory capable) epilogue that a stage's last conv uses when its output is the next GPU's slot
    DEFER_CHECK(fmt != DEFER_FMT_F32 && !x_is_f32, "k_conv: tcgen05 backend needs BF16X2/BF16 activations");
    DEFER_CHECK(umma_conv_supported(fmt, n, h, w, cin, ho, wo, cout, kh, kw, sh, sw, pad_t, pad_l),
                "k_conv: shape not supported by the tcgen05 kernel");
    UmmaConvPlan plan;
    UmmaConvLaneArgs args;
    void* dev_op = nullptr;
    const int want_bn = (backend & 1) ? 128 : 64;
    int rc = umma_conv_prepare(&plan, fmt, n, h, w, cin, ho, wo, cout, kh, kw, sh, sw, pad_t, pad_l, flags, w_hwio, scale, shift,
                               /*mega=*/false, /*stream_bn=*/want_bn);
    args.direct_out = backend >= 6;
    if (rc == DEFER_OK) rc = umma_conv_bind(plan, &args, x, residual, y);
    long long* trace = nullptr;
    const char* trace_path = getenv("DEFER_UMMA_TRACE");
    if (rc == DEFER_OK && trace_path) {
      cudaMalloc((void**)&trace, 8 * 64 * sizeof(long long));
      cudaMemset(trace, 0, 8 * 64 * sizeof(long long));
      args.trace = trace;
    }
    std::vector<unsigned char> host(umma_mega_op_bytes());
    if (rc == DEFER_OK) rc = umma_mega_fill(host.data(), plan, args);
    if (rc == DEFER_OK && cudaMalloc(&dev_op, host.size()) != cudaSuccess) rc = DEFER_ERR_CUDA;
    if (rc == DEFER_OK) cudaMemcpy(dev_op, host.data(), host.size(), cudaMemcpyHostToDevice);
    const int n_tiles = plan.tiles_n * plan.tiles_h * plan.tiles_w * (cout / plan.bn);
    if (rc == DEFER_OK && trace) {   // warm-up so the traced launch sees warm descriptor / instruction caches
      rc = launch_conv_stream(plan.nplanes, plan.bn, dev_op, n_tiles, plan.k_blocks, st);
      cudaStreamSynchronize(st);
      cudaMemset(trace, 0, 8 * 64 * sizeof(long long));
    }
    if (rc == DEFER_OK) rc = launch_conv_stream(plan.nplanes, plan.bn, dev_op, n_tiles, plan.k_blocks, st);
    cudaError_t e = cudaStreamSynchronize(st);
    if (trace) {
      long long hb[8 * 64];
      cudaMemcpy(hb, trace, sizeof hb, cudaMemcpyDeviceToHost);
      FILE* f = fopen(trace_path, "a");
      if (f) {
        fprintf(f, "# stream conv backend=%d n=%d hw=%dx%d cin=%d cout=%d k=%d s=%d bn=%d tiles=%d k_blocks=%d (ns since the first stamp)\n", backend,
                n, h, w, cin, cout, kh, sh, plan.bn, n_tiles, plan.k_blocks);
        for (int i = 0; i < 8; ++i) {
          if (!hb[i * 64]) continue;
          fprintf(f, "tile %d:", i);
          for (int j = 0; j < 64; ++j)
            if (hb[i * 64 + j]) fprintf(f, " %d=%lld", j, hb[i * 64 + j] - hb[0]);
          fprintf(f, "\n");
        }
        fclose(f);
      }
      cudaFree(trace);
    }
    if (dev_op) cudaFree(dev_op);
    umma_conv_unbind(&args);
    umma_conv_release(plan);
    if (rc != DEFER_OK) return rc;
    DEFER_CUDA(e);
    return DEFER_OK;
  }
  if (backend == 3) {
    // persistent-grid tcgen05 kernel (the mode the stage runtime picks for ops with many tiles)
    DEFER_CHECK(fmt != DEFER_FMT_F32 && !x_is_f32, "k_conv: tcgen05 backend needs BF16X2/BF16 activations");
    DEFER_CHECK(umma_conv_supported(fmt, n, h, w, cin, ho, wo, cout, kh, kw, sh, sw, pad_t, pad_l),
                "k_conv: shape not supported by the tcgen05 kernel");
    UmmaConvPlan plan;
    UmmaConvLaneArgs args;
    void* dev_op = nullptr;
    long long* trace = nullptr;
    const char* trace_path = getenv("DEFER_UMMA_TRACE");
    int rc = umma_conv_prepare(&plan, fmt, n, h, w, cin, ho, wo, cout, kh, kw, sh, sw, pad_t, pad_l, flags, w_hwio, scale, shift,
                               /*mega=*/true);
    if (rc == DEFER_OK) rc = umma_conv_bind(plan, &args, x, residual, y);
    if (rc == DEFER_OK && trace_path) {
      cudaMalloc((void**)&trace, 24 * 8 * sizeof(long long));
      cudaMemset(trace, 0, 24 * 8 * sizeof(long long));
      args.trace = trace;
    }
    std::vector<unsigned char> host(umma_mega_op_bytes());
    if (rc == DEFER_OK) rc = umma_mega_fill(host.data(), plan, args);
    if (rc == DEFER_OK && cudaMalloc(&dev_op, host.size()) != cudaSuccess) rc = DEFER_ERR_CUDA;
    if (rc == DEFER_OK) cudaMemcpy(dev_op, host.data(), host.size(), cudaMemcpyHostToDevice);
    const int n_tiles = plan.tiles_n * plan.tiles_h * plan.tiles_w * (cout / 64);
    if (rc == DEFER_OK && trace) {   // warm-up so the traced launch sees warm descriptor / instruction caches
      rc = launch_conv_persistent(plan.nplanes, dev_op, n_tiles, st);
      cudaStreamSynchronize(st);
      cudaMemset(trace, 0, 24 * 8 * sizeof(long long));
    }
    if (rc == DEFER_OK) rc = launch_conv_persistent(plan.nplanes, dev_op, n_tiles, st);
    cudaError_t e = cudaStreamSynchronize(st);
    if (trace) {
      long long hb[24 * 8];
      cudaMemcpy(hb, trace, sizeof hb, cudaMemcpyDeviceToHost);
      FILE* f = fopen(trace_path, "a");
      if (f) {
        fprintf(f, "# persistent conv n=%d hw=%dx%d cin=%d cout=%d k=%d s=%d tiles=%d\n", n, h, w, cin, cout, kh, sh, n_tiles);
        for (int i = 0; i < 24; ++i) {
          if (!hb[i * 8]) continue;
          fprintf(f, "%d", i);
          for (int j = 0; j < 8; ++j) fprintf(f, " %lld", hb[i * 8 + j] ? hb[i * 8 + j] - hb[0] : -1);
          fprintf(f, "\n");
        }
        fclose(f);
      }
      cudaFree(trace);
    }
    if (dev_op) cudaFree(dev_op);
    umma_conv_unbind(&args);
    umma_conv_release(plan);
    if (rc != DEFER_OK) return rc;
    DEFER_CUDA(e);
    return DEFER_OK;
  }
  if (backend == 2) {
    DEFER_CHECK(fmt != DEFER_FMT_F32 && !x_is_f32, "k_conv: tcgen05 backend needs BF16X2/BF16 activations");
    DEFER_CHECK(umma_conv_supported(fmt, n, h, w, cin, ho, wo, cout, kh, kw, sh, sw, pad_t, pad_l),
                "k_conv: shape not supported by the tcgen05 kernel");
    UmmaConvPlan plan;
    UmmaConvLaneArgs args;
    int rc = umma_conv_prepare(&plan, fmt, n, h, w, cin, ho, wo, cout, kh, kw, sh, sw, pad_t, pad_l, flags, w_hwio, scale, shift);
    if (rc == DEFER_OK) rc = umma_conv_bind(plan, &args, x, residual, y);
    long long* trace = nullptr;
    int n_ctas = plan.tiles_n * plan.tiles_h * plan.tiles_w * (plan.cout / plan.bn) * plan.splits;
    const char* trace_path = getenv("DEFER_UMMA_TRACE");
    if (rc == DEFER_OK && trace_path) {
      cudaMalloc((void**)&trace, (size_t)n_ctas * 8 * sizeof(long long));
      cudaMemset(trace, 0, (size_t)n_ctas * 8 * sizeof(long long));
      args.trace = trace;
      rc = launch_conv_umma(plan, args, st);   // warm-up launch (descriptor / instruction caches)
      cudaStreamSynchronize(st);
    }
    if (rc == DEFER_OK) rc = launch_conv_umma(plan, args, st);
    cudaError_t e = cudaStreamSynchronize(st);
    if (trace) {
      long long* h = (long long*)malloc((size_t)n_ctas * 8 * sizeof(long long));
      cudaMemcpy(h, trace, (size_t)n_ctas * 8 * sizeof(long long), cudaMemcpyDeviceToHost);
      FILE* f = fopen(trace_path, "a");
      if (f) {
        fprintf(f, "# conv n=%d h=%d w=%d cin=%d cout=%d k=%d s=%d bn=%d splits=%d stages=%d ctas=%d\n", n, h, w, cin, cout, kh, sh,
                plan.bn, plan.splits, plan.stages, n_ctas);
        for (int i = 0; i < n_ctas; ++i) {
          long long* t = h + 8 * i;
          fprintf(f, "%d %lld %lld %lld %lld %lld %lld %lld\n", i, t[7], t[1] - t[0], t[2] - t[0], t[3] - t[0], t[4] - t[0],
                  t[5] - t[0], t[6] - t[0]);
        }
        fclose(f);
      }
      free(h);
      cudaFree(trace);
    }
    umma_conv_unbind(&args);
    umma_conv_release(plan);
    if (rc != DEFER_OK) return rc;
    DEFER_CUDA(e);
    return DEFER_OK;
  }
  ConvParams p;
  p.x = x; p.w = w_hwio; p.scale = scale; p.shift = shift; p.res = residual; p.y = y;
  p.n = n; p.h = h; p.w_in = w; p.cin = cin; p.ho = ho; p.wo = wo; p.cout = cout;
  p.kh = kh; p.kw = kw; p.sh = sh; p.sw = sw; p.pad_t = pad_t; p.pad_l = pad_l; p.flags = flags;
  return launch_conv_simt(fmt, x_is_f32 != 0, p, st);
}

int defer_k_maxpool(int fmt, const void* x, void* y, int n, int h, int w, int c, int ph, int pw, int sh, int sw, int pad_t,
                    int pad_l, int pad_b, int pad_r, void* stream) {
  DEFER_CHECK(x && y, "k_maxpool: null pointer");
  int ho = (h + pad_t + pad_b - ph) / sh + 1, wo = (w + pad_l + pad_r - pw) / sw + 1;
  return launch_maxpool(fmt, x, y, n, h, w, c, ph, pw, sh, sw, pad_t, pad_l, ho, wo, (cudaStream_t)stream);
}

int defer_k_gap(int fmt, const void* x, void* y, int n, int h, int w, int c, void* stream) {
  DEFER_CHECK(x && y, "k_gap: null pointer");
  return launch_gap(fmt, x, y, n, h, w, c, (cudaStream_t)stream);
}

int defer_k_dense(int fmt, const void* x, const float* w_io, const float* bias, void* y, int y_is_f32, int n,
                  int in_features, int units, uint32_t flags, void* stream) {
  DEFER_CHECK(x && w_io && y, "k_dense: null pointer");
  float* partial = nullptr;
  size_t bytes = dense_workspace_bytes(n, in_features, units);
  DEFER_CUDA(cudaMalloc((void**)&partial, bytes));
  DEFER_CUDA(cudaMemsetAsync(partial, 0, bytes, (cudaStream_t)stream));
  int rc = launch_dense(fmt, x, w_io, false, bias, y, y_is_f32 != 0, partial, n, in_features, units, flags, (cudaStream_t)stream);
  cudaError_t e = cudaStreamSynchronize((cudaStream_t)stream);
  cudaFree(partial);
  if (rc != DEFER_OK) return rc;
  DEFER_CUDA(e);
  return DEFER_OK;
}

int defer_k_softmax(const float* x, float* y, int n, int c, void* stream) {
  DEFER_CHECK(x && y, "k_softmax: null pointer");
  return launch_softmax(x, y, n, c, (cudaStream_t)stream);
}

int defer_k_eltwise(int fmt, int kind, const void* a, const void* b, const float* scale, const float* shift, void* y, int n,
                    int h, int w, int c, uint32_t flags, void* stream) {
  DEFER_CHECK(a && y, "k_eltwise: null pointer");
  DEFER_CHECK(kind != DEFER_OP_ADD || b, "k_eltwise: ADD needs b");
  return launch_eltwise(fmt, kind, a, b, scale, shift, y, (size_t)n * h * w, c, flags, (cudaStream_t)stream);
}

int defer_k_encode(int fmt, const float* x_f32, void* y_act, uint64_t n_elems, void* stream) {
  DEFER_CHECK(x_f32 && y_act, "k_encode: null pointer");
  return launch_encode(fmt, x_f32, y_act, n_elems, (cudaStream_t)stream);
}

int defer_k_decode(int fmt, const void* x_act, float* y_f32, uint64_t n_elems, void* stream) {
  DEFER_CHECK(x_act && y_f32, "k_decode: null pointer");
  return launch_decode(fmt, x_act, y_f32, n_elems, (cudaStream_t)stream);
}

}  // extern "C"
