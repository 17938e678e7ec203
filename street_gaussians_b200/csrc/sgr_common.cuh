// sgr_common.cuh — shared device helpers, state layout and the exact tile-cull test.
//
// Data layout in HBM (all caller-owned, see include/sgr.h):
//   geom state   : GaussRec[P] (48 B packed record, 3 x float4 — one 16-B aligned gather of 3 vectors per splat
//                  instance in the blend passes), then tiles_touched u32[P], offsets u32[P], scan temp.
//   img state    : ranges uint2[Ntile], n_contrib u32[H*W], block-max n_contrib u32[Ntile]
//   binning state: keys_in u64[R], keys_out u64[R], vals_in u32[R], vals_out u32[R] (= tile-ordered point list), sort temp
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/sgr.h"

#define SGR_TILE 16            // tile edge in pixels (reference BLOCK_X = BLOCK_Y = 16, DGR/cuda_rasterizer/config.h:17-18)
#define SGR_TILE_PIX 256
#define SGR_ALIGN 256

namespace sgr {

// 48-byte per-Gaussian record written by preprocess_fwd and gathered by both blend passes.
//   q0 = (pix.x, pix.y, conic.xx, conic.xy)   q1 = (conic.yy, opacity, view depth, r)   q2 = (g, b, bits(clamped), unused)
struct __align__(16) GaussRec {
	float4 q0, q1, q2;
};

struct GeomView {
	GaussRec *rec;
	uint32_t *tiles_touched;
	uint32_t *offsets;
	void *scan_temp;
	size_t scan_temp_bytes;
	size_t total_bytes;
};
struct ImgView {
	uint2 *ranges;
	uint32_t *n_contrib;
	uint32_t *tile_max_contrib;
	size_t total_bytes;
};
struct BinView {
	uint64_t *keys_in, *keys_out;
	uint32_t *vals_in, *vals_out;
	void *sort_temp;
	size_t sort_temp_bytes;
	size_t total_bytes;
};

static inline size_t align_up(size_t v, size_t a = SGR_ALIGN) { return (v + a - 1) / a * a; }

// Column-major 3x3 (m[c][r]) with the product written as a left-associated sum of three products.  The reference
// uses glm::mat3, whose operator* has this exact algebraic form; keeping the form lets nvcc contract mul+add
// pairs identically so that depth / conic / radius come out bit-equal to the reference's (SURVEY.md §7
// "Bit-level parity traps").
struct M3 {
	float m[3][3];
};
__device__ __forceinline__ M3 m3_mul(const M3 &a, const M3 &b) {
	M3 r;
#pragma unroll
	for (int c = 0; c < 3; c++)
#pragma unroll
		for (int i = 0; i < 3; i++) r.m[c][i] = a.m[0][i] * b.m[c][0] + a.m[1][i] * b.m[c][1] + a.m[2][i] * b.m[c][2];
	return r;
}
__device__ __forceinline__ M3 m3_t(const M3 &a) {
	M3 r;
#pragma unroll
	for (int c = 0; c < 3; c++)
#pragma unroll
		for (int i = 0; i < 3; i++) r.m[c][i] = a.m[i][c];
	return r;
}

// Column-major affine / projective point transforms (reference: transformPoint4x3 / 4x4, auxiliary.h:58-77).
__device__ __forceinline__ float3 xform4x3(const float3 p, const float *__restrict__ m) {
	return make_float3(m[0] * p.x + m[4] * p.y + m[8] * p.z + m[12], m[1] * p.x + m[5] * p.y + m[9] * p.z + m[13],
	                   m[2] * p.x + m[6] * p.y + m[10] * p.z + m[14]);
}
__device__ __forceinline__ float4 xform4x4(const float3 p, const float *__restrict__ m) {
	return make_float4(m[0] * p.x + m[4] * p.y + m[8] * p.z + m[12], m[1] * p.x + m[5] * p.y + m[9] * p.z + m[13],
	                   m[2] * p.x + m[6] * p.y + m[10] * p.z + m[14], m[3] * p.x + m[7] * p.y + m[11] * p.z + m[15]);
}

// NDC -> pixel.  The reference writes the constants as double literals (auxiliary.h:41-44), i.e. fp64 arithmetic.
__device__ __forceinline__ float ndc2pix(float v, int S) { return ((v + 1.0) * S - 1.0) * 0.5; }

// 3-sigma tile rectangle, truncating toward zero before the clamp exactly like auxiliary.h:46-56.
__device__ __forceinline__ void tile_rect(float px, float py, int radius, int gx, int gy, int &x0, int &y0, int &x1, int &y1) {
	x0 = min(gx, max(0, (int)((px - radius) / SGR_TILE)));
	y0 = min(gy, max(0, (int)((py - radius) / SGR_TILE)));
	x1 = min(gx, max(0, (int)((px + radius + SGR_TILE - 1) / SGR_TILE)));
	y1 = min(gy, max(0, (int)((py + radius + SGR_TILE - 1) / SGR_TILE)));
}

// Which tile rows does this process own (multi-GPU tile-row sharding)?
struct Band {
	int begin, end, step;
};
__host__ __device__ __forceinline__ bool band_owns(const Band b, int row) {
	return row >= b.begin && row < b.end && ((row - b.begin) % b.step) == 0;
}
// dense index of an owned row within the band (used to launch one CTA per owned tile)
__host__ __device__ __forceinline__ int band_rows(const Band b) { return b.end > b.begin ? (b.end - b.begin + b.step - 1) / b.step : 0; }

// ---- exact, opacity-aware tile culling -------------------------------------------------------------------------
// A (Gaussian, tile) pair of the reference's 3-sigma rectangle can be dropped without changing ANY output iff no pixel
// of the tile passes the reference's two per-pixel tests (forward.cu:420-430):  power <= 0  and  o*exp(power) >= 1/255.
// With q(d) = a dx^2 + 2b dx dy + c dy^2 = -2*power this is  q <= 2*ln(255*o).  For a positive-definite conic the
// minimum of q over the tile's pixel rectangle is attained on an edge facing the centre, in closed form.  Everything
// is evaluated conservatively (slack far above fp32 rounding, "keep" on NaN / non-PD input).
struct CullParams {
	float mx, my, a, b, c, qmax;  // qmax < 0 => nothing can pass; qmax = +inf => keep all of the rectangle
};
__device__ __forceinline__ CullParams make_cull(float mx, float my, float a, float b, float c, float opacity) {
	CullParams cp{mx, my, a, b, c, __int_as_float(0x7f800000)};
	const bool pd = (a > 0.f) && (c > 0.f) && (a * c - b * b > 0.f);
	const float tau = logf(255.0f * opacity);  // NaN for negative / NaN opacity -> comparison below fails -> keep
	if (pd && tau == tau) cp.qmax = 2.0f * tau + (0.02f + 1e-3f * fabsf(tau));
	return cp;
}
__device__ __forceinline__ bool tile_can_contribute(const CullParams cp, int tx, int ty) {
	// pixel centres of the tile are the integers [x0, x0+15] x [y0, y0+15]; u = pixel - mean
	const float ux0 = (float)(tx * SGR_TILE) - cp.mx, ux1 = ux0 + (SGR_TILE - 1);
	const float uy0 = (float)(ty * SGR_TILE) - cp.my, uy1 = uy0 + (SGR_TILE - 1);
	const bool in_x = (ux0 <= 0.f) && (ux1 >= 0.f), in_y = (uy0 <= 0.f) && (uy1 >= 0.f);
	float qmin = 0.f;
	if (!(in_x && in_y)) {
		qmin = __int_as_float(0x7f800000);
		if (!in_x) {  // facing vertical edge
			const float ux = ux0 > 0.f ? ux0 : ux1;
			const float uy = fminf(fmaxf(-cp.b * ux / cp.c, uy0), uy1);
			qmin = cp.a * ux * ux + 2.f * cp.b * ux * uy + cp.c * uy * uy;
		}
		if (!in_y) {  // facing horizontal edge
			const float uy = uy0 > 0.f ? uy0 : uy1;
			const float ux = fminf(fmaxf(-cp.b * uy / cp.a, ux0), ux1);
			qmin = fminf(qmin, cp.a * ux * ux + 2.f * cp.b * ux * uy + cp.c * uy * uy);
		}
	}
	return !(qmin > cp.qmax);
}

// state carving (host) — implemented in capi.cu
GeomView carve_geom(void *base, int P);
ImgView carve_img(void *base, int W, int H);
BinView carve_bin(void *base, int64_t R);

// kernel launchers (host) — one per translation unit
struct FrameDev {  // SgrFrame + derived values, passed by value to kernels
	int P, D, M, S, W, H, gx, gy;
	float tanx, tany, fx, fy, mod;
	Band band;
	const float *bg, *view, *proj, *campos;
};

cudaError_t launch_preprocess_fwd(const FrameDev &f, const float *means3D, const float *shs, const float *colors_precomp,
                                  const float *opacities, const float *scales, const float *rotations,
                                  const float *cov3D_precomp, int32_t *radii, GeomView g, cudaStream_t st);
cudaError_t launch_filter(const FrameDev &f, const float *means3D, const float *scales, const float *rotations,
                          const float *cov3D_precomp, int32_t *radii, float *means2D, cudaStream_t st);
cudaError_t launch_mark_visible(int P, const float *means3D, const float *view, uint8_t *present, cudaStream_t st);
cudaError_t launch_scan(const FrameDev &f, GeomView g, cudaStream_t st);
size_t scan_temp_bytes(int P);
size_t sort_temp_bytes(int64_t R);
cudaError_t launch_binning(const FrameDev &f, GeomView g, const int32_t *radii, BinView b, ImgView img, int64_t R,
                           cudaStream_t st);
cudaError_t launch_blend_fwd(const FrameDev &f, GeomView g, BinView b, ImgView img, const float *semantics, float *out_color,
                             float *out_depth, float *out_alpha, float *out_sem, cudaStream_t st);
cudaError_t launch_blend_bwd(const FrameDev &f, GeomView g, BinView b, ImgView img, const float *semantics,
                             const float *out_alpha, const float *dL_dcolor, const float *dL_ddepth, const float *dL_dalpha,
                             const float *dL_dsem, float *grad2d, float *dL_dsemantics, cudaStream_t st);
cudaError_t launch_preprocess_bwd(const FrameDev &f, const float *means3D, const float *shs, const float *colors_precomp,
                                  const float *scales, const float *rotations, const float *cov3D_precomp,
                                  const int32_t *radii, GeomView g, const float *grad2d, float *dL_dmeans3D,
                                  float *dL_dmeans2D, float *dL_dsh, float *dL_dcolors, float *dL_dopacity,
                                  float *dL_dscales, float *dL_drot, float *dL_dcov3D, cudaStream_t st);
size_t knn_scratch_bytes(int P);
cudaError_t launch_knn(int P, const float *points, float *out, void *scratch, size_t scratch_bytes, cudaStream_t st);

}  // namespace sgr
