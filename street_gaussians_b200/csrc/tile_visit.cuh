// tile_visit.cuh — enumerate the tiles of one Gaussian's 3-sigma rectangle that (a) lie in this process's tile-row
// band and (b) pass the exact contribution test; either count them (preprocess) or emit (key, value) pairs (binning).
//
// Count and emit run the SAME code on the SAME inputs, so the per-Gaussian counts always match the offsets.
// Small rectangles are walked by the owning thread; rectangles larger than kCoopArea tiles are walked by the whole
// warp (lane-strided) so that one screen-filling splat cannot serialise 10^4 iterations on a single thread (the
// reference's duplicateWithKeys does exactly that, rasterizer_impl.cu:70-111).
//
// Key = tile_id << 32 | float_bits(view depth): identical to the reference (rasterizer_impl.cu:100-106).  Within one
// Gaussian the emission order is irrelevant (all its keys differ in the tile field); across Gaussians every instance
// of Gaussian i lands in [offset(i), offset(i)+count(i)), so a stable sort breaks depth ties by Gaussian index exactly
// like the reference.
#pragma once
#include "sgr_common.cuh"

namespace sgr {

constexpr int kCoopArea = 48;

template <bool EMIT>
__device__ __forceinline__ void visit_tiles(bool active, int x0, int y0, int x1, int y1, const CullParams cp, const Band band,
                                            int gx, uint32_t depth_bits, uint32_t gauss_idx, uint32_t offset,
                                            uint64_t *__restrict__ keys, uint32_t *__restrict__ vals, uint32_t &count) {
	const unsigned full = 0xffffffffu;
	const int lane = threadIdx.x & 31;
	// restrict the row range to the band up front (cheap and exact for contiguous bands)
	if (active) {
		y0 = max(y0, band.begin);
		y1 = min(y1, band.end);
		if (y1 <= y0 || cp.qmax < 0.f) active = false;
	}
	const int w = active ? (x1 - x0) : 0, h = active ? (y1 - y0) : 0;
	const int area = w * h;
	const bool coop = area > kCoopArea;
	count = 0;
	if (active && !coop) {
		uint32_t off = offset;
		for (int ty = y0; ty < y1; ty++) {
			if (band.step != 1 && !band_owns(band, ty)) continue;
			for (int tx = x0; tx < x1; tx++) {
				if (!tile_can_contribute(cp, tx, ty)) continue;
				if (EMIT) {
					keys[off] = ((uint64_t)(uint32_t)(ty * gx + tx) << 32) | depth_bits;
					vals[off] = gauss_idx;
				}
				off++;
			}
		}
		count = off - offset;
	}
	unsigned todo = __ballot_sync(full, coop);
	while (todo) {
		const int src = __ffs(todo) - 1;
		todo &= todo - 1;
		CullParams c2;
		c2.mx = __shfl_sync(full, cp.mx, src); c2.my = __shfl_sync(full, cp.my, src);
		c2.a = __shfl_sync(full, cp.a, src); c2.b = __shfl_sync(full, cp.b, src);
		c2.c = __shfl_sync(full, cp.c, src); c2.qmax = __shfl_sync(full, cp.qmax, src);
		const int sx0 = __shfl_sync(full, x0, src), sy0 = __shfl_sync(full, y0, src);
		const int sw = __shfl_sync(full, w, src), sarea = __shfl_sync(full, area, src);
		const uint32_t sdepth = __shfl_sync(full, depth_bits, src), sidx = __shfl_sync(full, gauss_idx, src);
		uint32_t base = __shfl_sync(full, offset, src);
		const uint32_t base0 = base;
		for (int t0 = 0; t0 < sarea; t0 += 32) {
			const int t = t0 + lane;
			bool keep = false;
			int tx = 0, ty = 0;
			if (t < sarea) {
				ty = sy0 + t / sw;
				tx = sx0 + t % sw;
				keep = band_owns(band, ty) && tile_can_contribute(c2, tx, ty);
			}
			const unsigned m = __ballot_sync(full, keep);
			if (EMIT && keep) {
				const uint32_t pos = base + __popc(m & ((1u << lane) - 1u));
				keys[pos] = ((uint64_t)(uint32_t)(ty * gx + tx) << 32) | sdepth;
				vals[pos] = sidx;
			}
			base += __popc(m);
		}
		if (lane == src) count = base - base0;
	}
}

}  // namespace sgr
