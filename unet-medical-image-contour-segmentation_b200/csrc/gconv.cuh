// Device-side view of unetb200_gconv_t and the dispatch hooks between the SIMT and tcgen05 engines.
#pragma once
#include "common.cuh"

namespace ub {

struct GconvDev {
  int B, Hm, Wm, Cin, ntaps;
  int tap_dy[9], tap_dx[9];
  int in_scale, in_off_y, in_off_x, Hin, Win;
  long long ld_in;
  int N, nquad, Cq, out_scale, out_off_y, out_off_x, Hout, Wout;
  long long ld_out;
  long long M;   // B*Hm*Wm
  int K;         // ntaps*Cin
};

int gconv_validate(const unetb200_gconv_t* d, GconvDev* out);

// tcgen05 engine (conv_tc.cu).  *_supported() return 1 when the shape/dtype/alignment fits.
int tc_fprop_supported(const unetb200_gconv_t* d, const void* x, const void* wp, const void* y);
long long tc_fprop_tiles(const unetb200_gconv_t* d);
int tc_fprop(const unetb200_gconv_t* d, const GconvDev& g, const void* x, const void* wp, const float* bias, void* y,
             double* stats, float* stats_ws, cudaStream_t stream);
int launch_stats_reduce(const float* ws, long long ntiles, int C2, double* stats, cudaStream_t s);
int tc_wgrad_supported(const unetb200_gconv_t* d, const void* x, const void* gy);
int tc_wgrad_splits(const unetb200_gconv_t* d, const GconvDev& g);
int tc_wgrad(const unetb200_gconv_t* d, const GconvDev& g, const void* x, const void* gy, float* partials, int splits,
             cudaStream_t stream);

// persistent tcgen05 engine, second generation (conv_tc2.cu)
int tc2_fprop_supported(const unetb200_gconv_t* d, const void* x, const void* wp, const void* y);
long long tc2_stats_workspace(const unetb200_gconv_t* d);
int tc2_fprop(const unetb200_gconv_t* d, const GconvDev& g, const void* x, const void* wp, const float* bias, void* y,
              double* stats, float* stats_ws, cudaStream_t stream);

int tc2_wgrad_supported(const unetb200_gconv_t* d, const void* x, const void* gy);
int tc2_wgrad_splits(const unetb200_gconv_t* d);
int tc2_wgrad(const unetb200_gconv_t* d, const GconvDev& g, const void* x, const void* gy, float* partials, int splits,
              cudaStream_t stream);

// CTA-pair tcgen05 engine for the 3x3 convolutions, third generation (conv_tc3.cu)
int tc3_fprop_supported(const unetb200_gconv_t* d, const void* x, const void* wp, const float* bias, const void* y);
long long tc3_stats_workspace(const unetb200_gconv_t* d);
int tc3_fprop(const unetb200_gconv_t* d, const GconvDev& g, const void* x, const void* wp, void* y, double* stats,
              float* stats_ws, cudaStream_t stream);

int tc3_wgrad_supported(const unetb200_gconv_t* d, const void* x, const void* gy);
int tc3_wgrad_splits(const unetb200_gconv_t* d);
int tc3_wgrad(const unetb200_gconv_t* d, const GconvDev& g, const void* x, const void* gy, float* partials, int splits,
              cudaStream_t stream);

// first-layer (C_in <= 4) CUDA-core kernels (conv_first.cu)
int first_fprop_supported(const unetb200_gconv_t* d, const void* y);
long long first_fprop_tiles(const unetb200_gconv_t* d);
int first_fprop(const unetb200_gconv_t* d, const GconvDev& g, const void* x, const void* wp, void* y, double* stats,
                float* stats_ws, cudaStream_t s);
int first_wgrad_supported(const unetb200_gconv_t* d, const void* gy);
int first_wgrad_splits(const unetb200_gconv_t* d);
int first_wgrad(const unetb200_gconv_t* d, const GconvDev& g, const void* x, const void* gy, float* partials,
                int splits, cudaStream_t s);

}  // namespace ub
