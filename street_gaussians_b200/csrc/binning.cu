// binning.cu — instance offsets (scan), (tile|depth, gaussian) pair emission, global radix sort, per-tile ranges.
//
// Replaces the reference's cub InclusiveSum + duplicateWithKeys + cub SortPairs + identifyTileRanges
// (DGR/cuda_rasterizer/rasterizer_impl.cu:278-321, 70-138).  Integer / bit work: results are exact and, given the same
// instance set, the tile-ordered list is identical to the reference's (stable LSD radix sort on the same key).
// The instance set itself is smaller than the reference's: tile_visit.cuh drops (Gaussian, tile) pairs that provably
// receive no contribution, which shrinks every HBM-bound pass in this file and both blend passes.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "sgr_common.cuh"
#include "tile_visit.cuh"

namespace sgr {

size_t scan_temp_bytes(int P) {
	size_t bytes = 0;
	cub::DeviceScan::InclusiveSum(nullptr, bytes, (uint32_t *)nullptr, (uint32_t *)nullptr, P > 0 ? P : 1);
	return bytes;
}
size_t sort_temp_bytes(int64_t R) {
	size_t bytes = 0;
	cub::DeviceRadixSort::SortPairs(nullptr, bytes, (uint64_t *)nullptr, (uint64_t *)nullptr, (uint32_t *)nullptr, (uint32_t *)nullptr,
	                                R > 0 ? R : 1);
	return bytes;
}

cudaError_t launch_scan(const FrameDev &f, GeomView g, cudaStream_t st) {
	if (f.P == 0) return cudaSuccess;
	size_t bytes = g.scan_temp_bytes;
	return cub::DeviceScan::InclusiveSum(g.scan_temp, bytes, g.tiles_touched, g.offsets, f.P, st);
}

__global__ void __launch_bounds__(256) emit_pairs_kernel(const FrameDev f, const GaussRec *__restrict__ rec, const int32_t *__restrict__ radii,
                                                        const uint32_t *__restrict__ tiles_touched, const uint32_t *__restrict__ offsets,
                                                        uint64_t *__restrict__ keys, uint32_t *__restrict__ vals) {
	const int idx = blockIdx.x * blockDim.x + threadIdx.x;
	bool active = false;
	int x0 = 0, y0 = 0, x1 = 0, y1 = 0;
	CullParams cp = {};
	uint32_t depth_bits = 0, off = 0;
	if (idx < f.P && tiles_touched[idx] > 0) {
		const float4 q0 = rec[idx].q0, q1 = rec[idx].q1;
		tile_rect(q0.x, q0.y, radii[idx], f.gx, f.gy, x0, y0, x1, y1);
		cp = make_cull(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y);
		depth_bits = __float_as_uint(q1.z);
		off = idx == 0 ? 0u : offsets[idx - 1];
		active = true;
	}
	uint32_t count;
	visit_tiles<true>(active, x0, y0, x1, y1, cp, f.band, f.gx, depth_bits, (uint32_t)idx, off, keys, vals, count);
}

// One thread per sorted instance: a tile's range starts / ends where the tile field of the key changes
// (reference identifyTileRanges, rasterizer_impl.cu:116-138).  ranges must be zero-initialised.
__global__ void __launch_bounds__(256) tile_ranges_kernel(int64_t L, const uint64_t *__restrict__ keys, uint2 *__restrict__ ranges) {
	const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= L) return;
	const uint32_t cur = (uint32_t)(keys[idx] >> 32);
	if (idx == 0)
		ranges[cur].x = 0;
	else {
		const uint32_t prev = (uint32_t)(keys[idx - 1] >> 32);
		if (cur != prev) {
			ranges[prev].y = (uint32_t)idx;
			ranges[cur].x = (uint32_t)idx;
		}
	}
	if (idx == L - 1) ranges[cur].y = (uint32_t)L;
}

static int bits_for(uint32_t n) {  // smallest b with (1 << b) > n - 1, i.e. enough bits for tile ids 0..n-1
	int b = 0;
	while (b < 32 && (1ull << b) < (unsigned long long)n) b++;
	return b;
}

cudaError_t launch_binning(const FrameDev &f, GeomView g, const int32_t *radii, BinView b, ImgView img, int64_t R, cudaStream_t st) {
	const int ntile = f.gx * f.gy;
	cudaError_t e = cudaMemsetAsync(img.ranges, 0, (size_t)ntile * sizeof(uint2), st);
	if (e != cudaSuccess) return e;
	if (R == 0 || f.P == 0) return cudaSuccess;
	emit_pairs_kernel<<<(f.P + 255) / 256, 256, 0, st>>>(f, g.rec, radii, g.tiles_touched, g.offsets, b.keys_in, b.vals_in);
	if ((e = cudaGetLastError()) != cudaSuccess) return e;
	size_t bytes = b.sort_temp_bytes;
	e = cub::DeviceRadixSort::SortPairs(b.sort_temp, bytes, b.keys_in, b.keys_out, b.vals_in, b.vals_out, R, 0, 32 + bits_for(ntile), st);
	if (e != cudaSuccess) return e;
	tile_ranges_kernel<<<(unsigned)((R + 255) / 256), 256, 0, st>>>(R, b.keys_out, img.ranges);
	return cudaGetLastError();
}

}  // namespace sgr
