// binning.cu — depth pre-sort, instance offsets, (tile, gaussian) pair emission, 2-pass stable tile sort, per-tile ranges.
//
// Replaces the reference's cub InclusiveSum + duplicateWithKeys + 64-bit cub SortPairs + identifyTileRanges
// (DGR/cuda_rasterizer/rasterizer_impl.cu:278-321, 70-138).  The reference sorts all R instances on a 46-bit
// (tile << 32 | depth bits) key: 6 onesweep passes x 24 B per instance.  Here the two halves of that key are sorted
// where they are cheapest:
//   1. the P Gaussians are sorted ONCE by their 32-bit depth bits (stable radix sort, ties keep ascending index);
//   2. instances are emitted in that order, so the unsorted list is already depth-ordered;
//   3. a STABLE radix sort on the tile id alone (<= 16 bits -> 2 passes, 16 B per instance per pass) then yields exactly
//      the reference's order: ascending tile, within a tile ascending depth bits, ties by ascending Gaussian index.
// Integer / bit work: the resulting per-tile lists equal the reference's lists minus the instances removed by the
// exact cull (tile_visit.cuh); tests/test_parity_gpu.py checks images bit-for-bit / to 1e-6 against the reference.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

#include "sgr_common.cuh"
#include "tile_visit.cuh"

namespace sgr {

struct GatherCount {
	const uint32_t *tiles_touched, *perm, *depth_sorted;
	// entries whose sorted key is 0xFFFFFFFF emit nothing: Gaussians without instances (uncompacted order) or the padding
	// behind the selected ones (compacted order, where perm is meaningless) — a real view depth > 0.2 never has those bits
	__host__ __device__ __forceinline__ uint32_t operator()(int t) const {
		return depth_sorted[t] == 0xffffffffu ? 0u : tiles_touched[perm[t]];
	}
};
struct WasDelivered {  // (predicate of the round-2 second-cut stream compaction; only sizes the temp storage now — the run tables replaced it)
	const int32_t *radii;
	__host__ __device__ __forceinline__ bool operator()(int i) const { return radii[i] > 0; }
};
using CountIter = cub::TransformInputIterator<uint32_t, GatherCount, cub::CountingInputIterator<int>>;

size_t geom_temp_bytes(int P) {
	const int n = P > 0 ? P : 1;
	size_t a = 0, b = 0;
	size_t c = 0;
	CountIter it(cub::CountingInputIterator<int>(0), GatherCount{nullptr, nullptr, nullptr});
	cub::DeviceScan::InclusiveSum(nullptr, a, it, (uint32_t *)nullptr, n);
	cub::DeviceRadixSort::SortPairs(nullptr, b, (uint32_t *)nullptr, (uint32_t *)nullptr, (uint32_t *)nullptr, (uint32_t *)nullptr, n);
	cub::DeviceSelect::If(nullptr, c, cub::CountingInputIterator<int>(0), (uint32_t *)nullptr, (uint32_t *)nullptr, n, WasDelivered{nullptr});
	a = a > b ? a : b;
	return a > c ? a : c;
}
size_t sort_temp_bytes(int64_t R) {
	size_t a = 0, b = 0;
	const int64_t n = R > 0 ? R : 1;
	cub::DeviceRadixSort::SortPairs(nullptr, a, (uint32_t *)nullptr, (uint32_t *)nullptr, (uint32_t *)nullptr, (uint32_t *)nullptr, n);
	cub::DeviceRadixSort::SortPairs(nullptr, b, (uint16_t *)nullptr, (uint16_t *)nullptr, (uint32_t *)nullptr, (uint32_t *)nullptr, n);
	return a > b ? a : b;
}

// Gaussian-sharded mode (include/sgr.h, sgr_forward_records): the records were projected on other ranks and gathered, so
// the tail of preprocess_fwd_kernel — instance count against THIS rank's tile-row band, depth sort key, identity
// permutation — runs here from the 48-B record + radius.  Same tile_rect / make_cull / visit_tiles as the emission
// kernels below, so counts and emitted ranges agree by construction.
// zero_rows (nullable): this rank's partial grad2d[P,12].  The rows of the Gaussians delivered to this rank (radius != 0 <=>
// rectangle meets the band) are the only ones its blend_bwd can touch and the only ones their owners read back, so they are
// zeroed here — 48 B x (Gaussians in the band) instead of a 48 B x P_total memset in front of every backward.
__global__ void __launch_bounds__(256) count_tiles_kernel(const FrameDev f, const GaussRec *__restrict__ rec,
                                                         const int32_t *__restrict__ radii, uint32_t *__restrict__ tiles_touched,
                                                         uint32_t *__restrict__ depth_key, uint32_t *__restrict__ iota,
                                                         float4 *__restrict__ zero_rows) {
	const int idx = blockIdx.x * blockDim.x + threadIdx.x;
	const bool in_range = idx < f.P;
	bool ok = false;
	int x0 = 0, y0 = 0, x1 = 0, y1 = 0;
	CullParams cp = {};
	float depth = 0.f;
	if (in_range) {
		const int r = radii[idx];
		if (r > 0) {
			const float4 q0 = rec[idx].q0, q1 = rec[idx].q1;
			tile_rect(q0.x, q0.y, r, f.gx, f.gy, x0, y0, x1, y1);
			cp = make_cull(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y);
			depth = q1.w;
			ok = true;
			if (zero_rows) {
				const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
				zero_rows[3 * (size_t)idx] = z; zero_rows[3 * (size_t)idx + 1] = z; zero_rows[3 * (size_t)idx + 2] = z;
			}
		}
	}
	uint32_t count = 0;
	visit_tiles<false, uint32_t>(ok, x0, y0, x1, y1, cp, f.band, f.gx, 0u, 0u, nullptr, nullptr, count);
	if (in_range) {
		tiles_touched[idx] = count;
		depth_key[idx] = count > 0 ? __float_as_uint(depth) : 0xffffffffu;
		iota[idx] = (uint32_t)idx;
	}
}
cudaError_t launch_count_tiles(const FrameDev &f, GeomView g, const int32_t *radii, cudaStream_t st, float *zero_rows) {
	if (f.P == 0) return cudaSuccess;
	count_launch();
	count_tiles_kernel<<<(f.P + 255) / 256, 256, 0, st>>>(f, g.rec, radii, g.tiles_touched, g.depth_key, g.iota,
	                                                      reinterpret_cast<float4 *>(zero_rows));
	return cudaGetLastError();
}

// depth order of the Gaussians + inclusive scan of their instance counts in that order (uncompacted: all f.P Gaussians take part,
// those without instances carry the key 0xFFFFFFFF and sort to the end).  tiles_touched / depth_key / iota come from
// preprocess_fwd_kernel<COUNT> or count_tiles_kernel.
cudaError_t launch_depth_order(const FrameDev &f, GeomView g, cudaStream_t st) {
	if (f.P == 0) return cudaSuccess;
	size_t bytes = g.temp_bytes;
	cudaError_t e = cub::DeviceRadixSort::SortPairs(g.temp, bytes, g.depth_key, g.depth_sorted, g.iota, g.perm, f.P, 0, 32, st);
	if (e != cudaSuccess) return e;
	CountIter it(cub::CountingInputIterator<int>(0), GatherCount{g.tiles_touched, g.perm, g.depth_sorted});
	bytes = g.temp_bytes;
	return cub::DeviceScan::InclusiveSum(g.temp, bytes, it, g.offsets, f.P, st);
}

// ---- fused Gaussian-sharded forward: the delivered runs (sgr_common.cuh "block-run exchange") -> compact depth-sort input ----
// cnt[s * nblk + b] = number of records block b of owner s delivered to this rank (written by the owners' projection kernels before the
// barrier).  One block turns the table into an exclusive prefix (entry e = first compact position of run e) and the total (status[4]).
__global__ void __launch_bounds__(1024) run_prefix_kernel(const uint32_t *__restrict__ cnt, uint32_t *__restrict__ pre, const uint32_t entries,
                                                          uint32_t *__restrict__ status) {
	__shared__ uint32_t s_warp[32];
	__shared__ uint32_t s_carry;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	if (threadIdx.x == 0) s_carry = 0u;
	__syncthreads();
	for (uint32_t base = 0; base < entries; base += 1024u) {
		const uint32_t e = base + threadIdx.x;
		const uint32_t v = e < entries ? cnt[e] : 0u;
		uint32_t incl = v;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
			if (lane >= o) incl += t;
		}
		if (lane == 31) s_warp[warp] = incl;
		__syncthreads();
		if (warp == 0) {
			uint32_t w = s_warp[lane];
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) {
				const uint32_t t = __shfl_up_sync(0xffffffffu, w, o);
				if (lane >= o) w += t;
			}
			s_warp[lane] = w;  // inclusive prefix of the warp totals
		}
		__syncthreads();
		const uint32_t carry = s_carry;
		const uint32_t excl = carry + (warp ? s_warp[warp - 1] : 0u) + incl - v;
		if (e < entries) pre[e] = excl;
		__syncthreads();
		if (threadIdx.x == 1023) s_carry = carry + s_warp[31];
		__syncthreads();
	}
	if (threadIdx.x == 0) status[4] = s_carry;
}

// thread j = j-th delivered Gaussian in ascending slot order (= ascending global id): locate its run by binary search in the prefix table,
// take the radius out of the record, count its tiles against this rank's band, write the depth-sort input and clear its grad2d row.
__global__ void __launch_bounds__(256) count_runs_kernel(const FrameDev f, const GaussRec *__restrict__ rec, int32_t *__restrict__ radii,
                                                        const uint32_t *__restrict__ pre, const uint32_t entries, const uint32_t nblk,
                                                        const long long chunk, uint32_t *__restrict__ status, const uint32_t cap_v,
                                                        uint32_t *__restrict__ tiles_touched, uint32_t *__restrict__ ckey,
                                                        uint32_t *__restrict__ cval, float4 *__restrict__ zero_rows) {
	const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t n_act = status[4];
	if (j == 0 && n_act > cap_v) atomicOr(&status[2], 2u);
	const bool live = j < n_act && j < cap_v;
	bool ok = false;
	int x0 = 0, y0 = 0, x1 = 0, y1 = 0;
	CullParams cp = {};
	float depth = 0.f;
	uint32_t gidx = 0u;
	if (live) {
		uint32_t lo = 0u, hi = entries;  // largest e with pre[e] <= j  (pre[0] = 0 <= j)
		while (hi - lo > 1u) {
			const uint32_t mid = (lo + hi) >> 1;
			if (pre[mid] <= j) lo = mid; else hi = mid;
		}
		const uint32_t s = lo / nblk, b = lo - s * nblk;
		gidx = (uint32_t)((long long)s * chunk + (long long)b * kRunBlock) + (j - pre[lo]);
		const float4 q0 = rec[gidx].q0, q1 = rec[gidx].q1;
		const int r = packed_radius(rec[gidx].q2.w);
		radii[gidx] = r;  // emit_pairs / emit_big read the radius from here
		tile_rect(q0.x, q0.y, r, f.gx, f.gy, x0, y0, x1, y1);
		cp = make_cull(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y);
		depth = q1.w;
		ok = true;
		if (zero_rows) {
			const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
			zero_rows[3 * (size_t)gidx] = z; zero_rows[3 * (size_t)gidx + 1] = z; zero_rows[3 * (size_t)gidx + 2] = z;
		}
	}
	uint32_t count = 0;
	visit_tiles<false, uint32_t>(ok, x0, y0, x1, y1, cp, f.band, f.gx, 0u, 0u, nullptr, nullptr, count);
	if (j < cap_v) {
		ckey[j] = (live && count > 0u) ? __float_as_uint(depth) : 0xffffffffu;
		cval[j] = gidx;
		if (live) tiles_touched[gidx] = count;
	}
}

// g.iota = count table (filled by the owners), g.perm = its prefix (dead before the sort writes g.perm), cap_v < 0 -> all f.P slots
cudaError_t launch_count_and_order_runs(const FrameDev &f, GeomView g, int32_t *radii, int world, long long chunk, cudaStream_t st, int64_t cap_v,
                                        float *zero_rows, int *n_order) {
	if (n_order) *n_order = f.P;
	if (f.P == 0) return cudaSuccess;
	const int n = (int)(cap_v < 0 ? (int64_t)f.P : (cap_v < 1 ? 1 : (cap_v > f.P ? f.P : cap_v)));
	const uint32_t nblk = (uint32_t)((chunk + kRunBlock - 1) / kRunBlock);
	const uint32_t entries = (uint32_t)world * nblk;
	// the prefix table lives in g.offsets until the count kernel has consumed it (the scan below then overwrites it)
	uint32_t *pre = g.offsets;
	count_launch();
	run_prefix_kernel<<<1, 1024, 0, st>>>(g.iota, pre, entries, g.big_count);
	count_launch();
	count_runs_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(f, g.rec, radii, pre, entries, nblk, chunk, g.big_count, (uint32_t)n, g.tiles_touched,
	                                                               g.ckey, g.cval, reinterpret_cast<float4 *>(zero_rows));
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) return e;
	size_t bytes = g.temp_bytes;
	e = cub::DeviceRadixSort::SortPairs(g.temp, bytes, g.ckey, g.depth_sorted, g.cval, g.perm, n, 0, 32, st);
	if (e != cudaSuccess) return e;
	if (n_order) *n_order = n;
	CountIter it(cub::CountingInputIterator<int>(0), GatherCount{g.tiles_touched, g.perm, g.depth_sorted});
	bytes = g.temp_bytes;
	return cub::DeviceScan::InclusiveSum(g.temp, bytes, it, g.offsets, n, st);
}

constexpr uint32_t kEmitStage = 512;  // instances staged per warp (2 x 2 KB of shared memory per warp)

// thread t handles the t-th Gaussian in depth order.  KeyT = uint16_t whenever the image has at most 65,536 tiles (up to
// 4096 x 4096 px): the tile sort then moves 6 instead of 8 bytes per pair per pass.
template <typename KeyT>
__global__ void __launch_bounds__(256) emit_pairs_kernel(const FrameDev f, const GaussRec *__restrict__ rec, const int32_t *__restrict__ radii,
                                                        const uint32_t *__restrict__ /*tiles_touched*/, const uint32_t *__restrict__ perm,
                                                        const uint32_t *__restrict__ offsets, KeyT *__restrict__ keys,
                                                        uint32_t *__restrict__ vals, uint32_t *__restrict__ big_list,
                                                        uint32_t *__restrict__ big_count, const uint32_t cap, const int n_order) {
	__shared__ KeyT s_keys[8][kEmitStage];
	__shared__ uint32_t s_vals[8][kEmitStage];
	const int warp = threadIdx.x >> 5;
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	bool active = false;
	int x0 = 0, y0 = 0, x1 = 0, y1 = 0;
	CullParams cp = {};
	uint32_t gidx = 0;
	// [off, end) = this Gaussian's output range (offsets = inclusive scan in depth order); lanes past P get an empty range
	const int tc = min(t, n_order - 1);  // n_order = depth-order slots: f.P, or the Gaussian capacity of the compacted mode
	uint32_t end = offsets[tc];
	uint32_t off = tc == 0 ? 0u : offsets[tc - 1];
	if (t >= n_order) off = end;
	// bounded mode: a Gaussian whose range does not fit the caller's capacity emits nothing.  BOTH ends are clamped so that
	// warp_first / warp_last (taken from lanes 0 / 31 below) never address past `cap`, even when every lane of the warp
	// overflows: the staged flush then covers at most [warp_first, cap), and pad_keys_kernel rewrites [emitted, cap).
	if (end > cap) {
		off = min(off, cap);
		end = off;
	}
	if (t < n_order) {
		if (end > off) {
			gidx = perm[t];  // (padding slots of the compacted order have end == off and a meaningless perm)
			const float4 q0 = rec[gidx].q0, q1 = rec[gidx].q1;
			tile_rect(q0.x, q0.y, radii[gidx], f.gx, f.gy, x0, y0, x1, y1);
			cp = make_cull(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y);
			active = true;
		}
	}
	const uint32_t warp_first = __shfl_sync(0xffffffffu, off, 0);
	const uint32_t warp_last = __shfl_sync(0xffffffffu, end, 31);
	const uint32_t warp_total = warp_last >= warp_first ? warp_last - warp_first : 0xffffffffu;  // (overflow tail: no staging)
	uint32_t count;
	visit_tiles<true, KeyT>(active, x0, y0, x1, y1, cp, f.band, f.gx, gidx, off, keys, vals, count, s_keys[warp], s_vals[warp], kEmitStage,
	                  warp_first, warp_total, big_list, big_count, (uint32_t)t);
}

// One warp per deferred (large-rectangle) Gaussian, grid-strided over the device-side list: no host round trip for the count.
template <typename KeyT>
__global__ void __launch_bounds__(256) emit_big_kernel(const FrameDev f, const GaussRec *__restrict__ rec, const int32_t *__restrict__ radii,
                                                      const uint32_t *__restrict__ perm, const uint32_t *__restrict__ offsets,
                                                      KeyT *__restrict__ keys, uint32_t *__restrict__ vals,
                                                      const uint32_t *__restrict__ big_list, const uint32_t *__restrict__ big_count,
                                                      const uint32_t cap) {
	const uint32_t n = *big_count;
	const uint32_t nwarps = gridDim.x * 8u;
	const int lane = threadIdx.x & 31;
	for (uint32_t i = blockIdx.x * 8u + (threadIdx.x >> 5); i < n; i += nwarps) {
		const uint32_t t = big_list[i];
		const uint32_t gidx = perm[t];
		const float4 q0 = rec[gidx].q0, q1 = rec[gidx].q1;
		int x0, y0, x1, y1;
		tile_rect(q0.x, q0.y, radii[gidx], f.gx, f.gy, x0, y0, x1, y1);
		const CullParams cp = make_cull(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y);
		const uint32_t off = t == 0 ? 0u : offsets[t - 1];
		if (offsets[t] > cap) continue;  // bounded mode overflow (never deferred in practice: emit_pairs already skipped it)
		// lane 0 carries the Gaussian through the cooperative path of visit_tiles (all other lanes inactive)
		uint32_t count;
		visit_tiles<true, KeyT>(lane == 0, x0, y0, x1, y1, cp, f.band, f.gx, gidx, off, keys, vals, count);
	}
}

// One thread per sorted instance: a tile's range starts / ends where the tile id changes
// (reference identifyTileRanges, rasterizer_impl.cu:116-138).  ranges must be zero-initialised.
constexpr int kRangeKeysPerThread = 8;
template <typename KeyT>
__global__ void __launch_bounds__(256) tile_ranges_kernel(int64_t L, const KeyT *__restrict__ keys, uint2 *__restrict__ ranges, const uint32_t ntile) {
	// 8 consecutive keys per thread (one 16-B load for 16-bit tile ids): the one-key-per-thread version ran at 0.6 TB/s.
	const int64_t base = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * kRangeKeysPerThread;
	if (base >= L) return;
	uint32_t k[kRangeKeysPerThread];
	if (base + kRangeKeysPerThread <= L) {
		if (sizeof(KeyT) == 2) {
			const uint4 v = *reinterpret_cast<const uint4 *>(keys + base);  // base is a multiple of 8 -> 16-B aligned
			k[0] = v.x & 0xffffu; k[1] = v.x >> 16; k[2] = v.y & 0xffffu; k[3] = v.y >> 16;
			k[4] = v.z & 0xffffu; k[5] = v.z >> 16; k[6] = v.w & 0xffffu; k[7] = v.w >> 16;
		} else {
			const uint4 a = reinterpret_cast<const uint4 *>(keys + base)[0], b = reinterpret_cast<const uint4 *>(keys + base)[1];
			k[0] = a.x; k[1] = a.y; k[2] = a.z; k[3] = a.w; k[4] = b.x; k[5] = b.y; k[6] = b.z; k[7] = b.w;
		}
	} else {
#pragma unroll
		for (int i = 0; i < kRangeKeysPerThread; i++) k[i] = base + i < L ? (uint32_t)keys[base + i] : 0xffffffffu;
	}
	// keys >= ntile are the padding of the bounded mode (they sort behind every real tile)
	uint32_t prev = base == 0 ? 0xffffffffu : (uint32_t)keys[base - 1];
#pragma unroll
	for (int i = 0; i < kRangeKeysPerThread; i++) {
		const int64_t idx = base + i;
		if (idx >= L) break;
		const uint32_t cur = k[i];
		if (cur != prev || idx == 0) {
			if (idx != 0 && prev < ntile) ranges[prev].y = (uint32_t)idx;
			if (cur < ntile) ranges[cur].x = (uint32_t)idx;
		}
		if (idx == L - 1 && cur < ntile) ranges[cur].y = (uint32_t)L;
		prev = cur;
	}
}

// status words: [1] = R (true instance count), [2] = overflow flag, [3] = instances actually emitted (<= cap)
__global__ void count_status_kernel(const uint32_t *__restrict__ offsets, int P, uint32_t cap, uint32_t *__restrict__ status) {
	const uint32_t R = offsets[P - 1];
	uint32_t emitted = R;
	if (R > cap) {  // largest scanned offset that still fits: Gaussians are emitted in depth order, the tail is dropped
		int lo = 0, hi = P;  // first index with offsets[i] > cap
		while (lo < hi) {
			const int mid = (lo + hi) >> 1;
			if (offsets[mid] > cap) hi = mid; else lo = mid + 1;
		}
		emitted = lo == 0 ? 0u : offsets[lo - 1];
	}
	status[1] = R;
	status[2] |= R > cap ? 1u : 0u;  // (bit 1 = Gaussian-capacity overflow of the compacted depth order; the words are zeroed per forward)
	status[3] = emitted;
}
// grid-stride over [emitted, cap): the number of padding slots is only known on the device, and one thread per slot of the whole
// capacity (the round-1 launch) spent 31 us retiring 55k empty blocks
template <typename KeyT>
__global__ void __launch_bounds__(256) pad_keys_kernel(KeyT *__restrict__ keys, uint32_t *__restrict__ vals, uint32_t cap,
                                                      const uint32_t *__restrict__ status) {
	const uint32_t stride = gridDim.x * blockDim.x;
	for (uint32_t i = status[3] + blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += stride) {
		keys[i] = (KeyT)~(KeyT)0;
		vals[i] = 0u;
	}
}

static int bits_for(uint32_t n) {  // smallest b with (1 << b) >= n, i.e. enough bits for tile ids 0..n-1
	int b = 0;
	while (b < 32 && (1ull << b) < (unsigned long long)n) b++;
	return b > 0 ? b : 1;
}

cudaError_t launch_binning(const FrameDev &f, GeomView g, const int32_t *radii, BinView b, ImgView img, int64_t R, cudaStream_t st,
                           int64_t cap, int n_order) {
	if (n_order < 0) n_order = f.P;
	const int ntile = f.gx * f.gy;
	const bool bounded = cap >= 0;
	cudaError_t e = cudaMemsetAsync(img.ranges, 0, (size_t)ntile * sizeof(uint2), st);
	if (e != cudaSuccess) return e;
	if (f.P == 0) return cudaSuccess;
	const uint32_t cap32 = bounded ? (uint32_t)cap : 0xffffffffu;
	count_launch();
	count_status_kernel<<<1, 1, 0, st>>>(g.offsets, n_order, cap32, g.big_count);
	if ((e = cudaGetLastError()) != cudaSuccess) return e;
	if (bounded) R = cap;  // every pass below runs over the caller's capacity; the padding carries the largest key
	if (R == 0) return cudaSuccess;
	size_t bytes = b.sort_temp_bytes;
	const unsigned nblk = (unsigned)((n_order + 255) / 256);
	const unsigned nbig_blk = 148 * 4;  // persistent-style grid for the deferred large rectangles
	if ((e = cudaMemsetAsync(g.big_count, 0, sizeof(uint32_t), st)) != cudaSuccess) return e;
	const unsigned npad_blk = bounded ? (unsigned)(cap / 256 + 1 < 148 * 8 ? cap / 256 + 1 : 148 * 8) : 0u;
	const int sort_bits = bounded ? (ntile <= 65535 ? 16 : 32) : bits_for(ntile);
	if (ntile <= 65535 || (!bounded && ntile <= 65536)) {  // 16-bit tile ids (the key arrays are allocated for 32-bit ids either way)
		uint16_t *kin = reinterpret_cast<uint16_t *>(b.keys_in), *kout = reinterpret_cast<uint16_t *>(b.keys_out);
		count_launch();
		emit_pairs_kernel<uint16_t><<<nblk, 256, 0, st>>>(f, g.rec, radii, g.tiles_touched, g.perm, g.offsets, kin, b.vals_in, g.big_list, g.big_count, cap32, n_order);
		count_launch();
		emit_big_kernel<uint16_t><<<nbig_blk, 256, 0, st>>>(f, g.rec, radii, g.perm, g.offsets, kin, b.vals_in, g.big_list, g.big_count, cap32);
		if (bounded) count_launch();
		if (bounded) pad_keys_kernel<uint16_t><<<npad_blk, 256, 0, st>>>(kin, b.vals_in, cap32, g.big_count);
		if ((e = cudaGetLastError()) != cudaSuccess) return e;
		e = cub::DeviceRadixSort::SortPairs(b.sort_temp, bytes, kin, kout, b.vals_in, b.vals_out, R, 0, sort_bits, st);
		if (e != cudaSuccess) return e;
		count_launch();
		tile_ranges_kernel<uint16_t><<<(unsigned)((R + 256 * kRangeKeysPerThread - 1) / (256 * kRangeKeysPerThread)), 256, 0, st>>>(R, kout, img.ranges, (uint32_t)ntile);
	} else {
		count_launch();
		emit_pairs_kernel<uint32_t><<<nblk, 256, 0, st>>>(f, g.rec, radii, g.tiles_touched, g.perm, g.offsets, b.keys_in, b.vals_in, g.big_list, g.big_count, cap32, n_order);
		count_launch();
		emit_big_kernel<uint32_t><<<nbig_blk, 256, 0, st>>>(f, g.rec, radii, g.perm, g.offsets, b.keys_in, b.vals_in, g.big_list, g.big_count, cap32);
		if (bounded) count_launch();
		if (bounded) pad_keys_kernel<uint32_t><<<npad_blk, 256, 0, st>>>(b.keys_in, b.vals_in, cap32, g.big_count);
		if ((e = cudaGetLastError()) != cudaSuccess) return e;
		e = cub::DeviceRadixSort::SortPairs(b.sort_temp, bytes, b.keys_in, b.keys_out, b.vals_in, b.vals_out, R, 0, sort_bits, st);
		if (e != cudaSuccess) return e;
		count_launch();
		tile_ranges_kernel<uint32_t><<<(unsigned)((R + 256 * kRangeKeysPerThread - 1) / (256 * kRangeKeysPerThread)), 256, 0, st>>>(R, b.keys_out, img.ranges, (uint32_t)ntile);
	}
	return cudaGetLastError();
}

}  // namespace sgr
