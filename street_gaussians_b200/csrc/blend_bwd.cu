// blend_bwd.cu — per-tile back-to-front backward blend.
//
// Replaces the reference's BACKWARD::renderCUDA<3,20> (DGR/cuda_rasterizer/backward.cu:415-641).  The per-(pixel, splat)
// arithmetic is the reference's (same recurrences for T, the "colour behind" accumulators and dL/dalpha); what changes
// is how the 11 per-Gaussian partial sums leave the SM.  The reference issues 11 scalar atomicAdd per contributing
// (pixel, splat) pair, all 256 threads of a tile hammering the same 44 bytes.  Here:
//   1. per splat, each warp (a compact 8x4 pixel block) reduces its 32 lanes with a TRANSPOSED butterfly: 16 values are
//      reduced with 8+4+2+1+1 = 16 shuffles (instead of 5 per value = 55), leaving component c on lanes 2c, 2c+1;
//      warps in which no lane contributes (ballot == 0) skip the reduction altogether;
//   2. the 8 warps park their partial sums in a private shared-memory slab s_part[warp][slot][12] (plain stores, no
//      atomics, no zero-fill: a 64-bit per-warp mask says which slots are live);
//   3. once per 64-splat batch, 4 threads per splat add up the live slabs and issue ONE global atomic per
//      (tile, splat, component): R*11 reductions per frame in total instead of (contributing pairs)*11.
// The list is walked from tile_max_contrib (deepest position any pixel of the tile reached in the forward pass), so the
// unreachable tail of a saturated tile's list is never touched.
#include "sgr_common.cuh"

namespace sgr {

constexpr int kBwdBatch = 64;
constexpr int kNComp = 12;  // 11 used + 1 pad (see sgr.h: grad2d layout)

// Transposed warp reduction of 16 values per lane: afterwards lane L holds the full 32-lane sum of component
// c(L) = 8*b4 + 4*b3 + 2*b2 + b1 (b_k = bit k of L) in v[0].
__device__ __forceinline__ float warp_reduce16_transposed(float (&v)[16], const int lane) {
	const unsigned full = 0xffffffffu;
	{
		const bool up = lane & 16;
#pragma unroll
		for (int i = 0; i < 8; i++) {
			const float send = up ? v[i] : v[i + 8];
			const float keep = up ? v[i + 8] : v[i];
			v[i] = keep + __shfl_xor_sync(full, send, 16);
		}
	}
	{
		const bool up = lane & 8;
#pragma unroll
		for (int i = 0; i < 4; i++) {
			const float send = up ? v[i] : v[i + 4];
			const float keep = up ? v[i + 4] : v[i];
			v[i] = keep + __shfl_xor_sync(full, send, 8);
		}
	}
	{
		const bool up = lane & 4;
#pragma unroll
		for (int i = 0; i < 2; i++) {
			const float send = up ? v[i] : v[i + 2];
			const float keep = up ? v[i + 2] : v[i];
			v[i] = keep + __shfl_xor_sync(full, send, 4);
		}
	}
	{
		const bool up = lane & 2;
		const float send = up ? v[0] : v[1];
		const float keep = up ? v[1] : v[0];
		v[0] = keep + __shfl_xor_sync(full, send, 2);
	}
	v[0] += __shfl_xor_sync(full, v[0], 1);
	return v[0];
}

template <int SCH>
__global__ void __launch_bounds__(256) blend_bwd_kernel(const FrameDev f, const uint2 *__restrict__ ranges,
                                                        const uint32_t *__restrict__ point_list, const GaussRec *__restrict__ rec,
                                                        const float *__restrict__ semantics, const uint32_t *__restrict__ n_contrib,
                                                        const uint32_t *__restrict__ tile_max_contrib, const float *__restrict__ alphas,
                                                        const float *__restrict__ dL_dpixels, const float *__restrict__ dL_dpixel_depths,
                                                        const float *__restrict__ dL_dalphas, const float *__restrict__ dL_dpixel_sems,
                                                        float *__restrict__ grad2d, float *__restrict__ dL_dsemantics) {
	__shared__ float4 s_q0[2][kBwdBatch];
	__shared__ float4 s_q1[2][kBwdBatch];
	__shared__ float4 s_q2[2][kBwdBatch];
	__shared__ uint32_t s_id[2][kBwdBatch];
	__shared__ float s_part[8][kBwdBatch][kNComp];
	__shared__ unsigned long long s_mask[8];

	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const int tile_x = blockIdx.x, tile_y = f.band.begin + blockIdx.y * f.band.step;
	const int tile = tile_y * f.gx + tile_x;
	const int n_eff = (int)tile_max_contrib[tile];
	if (n_eff == 0) return;
	const bool use_mask = f.P < (1 << 24);  // point-list value = warp mask << 24 | index (tile_visit.cuh)
	const uint32_t idx_mask = use_mask ? kIdxMask : 0xffffffffu;
	const uint32_t my_bit = use_mask ? (1u << (24 + warp)) : 0u;
	const int px = tile_x * SGR_TILE + (warp & 1) * 8 + (lane & 7);
	const int py = tile_y * SGR_TILE + (warp >> 1) * 4 + (lane >> 3);
	const bool inside = px < f.W && py < f.H;
	const size_t HW = (size_t)f.W * f.H;
	const size_t pix_id = (size_t)f.W * py + px;
	const float2 pixf = make_float2((float)px, (float)py);
	const uint32_t list0 = ranges[tile].x;
	const int nb = (n_eff + kBwdBatch - 1) / kBwdBatch;

	const float T_final = inside ? (1 - alphas[pix_id]) : 0;
	float T = T_final;
	const int last_contributor = inside ? (int)n_contrib[pix_id] : 0;
	float accum_rec[3] = {0, 0, 0}, last_color[3] = {0, 0, 0}, dL_dpixel[3] = {0, 0, 0};
	float accum_depth_rec = 0, last_depth = 0, accum_alpha_rec = 0, last_alpha = 0;
	float dL_dpixel_depth = 0, dL_dalpha_px = 0;
	float accum_sem[SCH > 0 ? SCH : 1], last_sem[SCH > 0 ? SCH : 1], dL_dsem_px[SCH > 0 ? SCH : 1];
#pragma unroll
	for (int c = 0; c < (SCH > 0 ? SCH : 1); c++) accum_sem[c] = last_sem[c] = dL_dsem_px[c] = 0.f;
	if (inside) {
#pragma unroll
		for (int c = 0; c < 3; c++) dL_dpixel[c] = dL_dpixels[c * HW + pix_id];
		dL_dpixel_depth = dL_dpixel_depths[pix_id];
		dL_dalpha_px = dL_dalphas[pix_id];
#pragma unroll
		for (int c = 0; c < SCH; c++)
			if (c < f.S) dL_dsem_px[c] = dL_dpixel_sems[c * HW + pix_id];
	}
	float bg_dot_dpixel = 0;
#pragma unroll
	for (int c = 0; c < 3; c++) bg_dot_dpixel += f.bg[c] * dL_dpixel[c];
	const float ddelx_dx = 0.5f * f.W, ddely_dy = 0.5f * f.H;

	// staging: 4 threads per record (q0, q1, q2, id); slot j of batch b <-> list index (n_eff - b*B) - 1 - j
	const int ld_slot = tid >> 2, ld_part = tid & 3;
	float4 rq = make_float4(0, 0, 0, 0);
	uint32_t rid = 0;
	auto fetch = [&](int b) {
		const int idx = n_eff - b * kBwdBatch - 1 - ld_slot;
		if (idx >= 0) {
			rid = point_list[list0 + idx];
			if (ld_part < 3) rq = reinterpret_cast<const float4 *>(rec + (rid & idx_mask))[ld_part];
		}
	};
	auto stash = [&](int buf) {
		if (ld_part == 0) s_q0[buf][ld_slot] = rq;
		else if (ld_part == 1) s_q1[buf][ld_slot] = rq;
		else if (ld_part == 2) s_q2[buf][ld_slot] = rq;
		else s_id[buf][ld_slot] = rid;
	};
	fetch(0);
	stash(0);

	for (int b = 0; b < nb; b++) {
		__syncthreads();  // buffer b&1 is published; s_part / s_mask of the previous batch have been consumed
		if (b + 1 < nb) fetch(b + 1);
		const int buf = b & 1;
		const int hi = n_eff - b * kBwdBatch;  // list position (1-based) of slot 0
		const int cnt = min(kBwdBatch, hi);
		unsigned long long wmask = 0ull;
		for (int j = 0; j < cnt; j++) {
			const int contributor = hi - 1 - j;  // 0-based list index of this slot
			if (my_bit != 0u && (s_id[buf][j] & my_bit) == 0u) continue;  // warp-uniform skip: block cannot receive anything
			bool valid = contributor < last_contributor;
			float v[16];
#pragma unroll
			for (int k = 0; k < 16; k++) v[k] = 0.f;
			float sem_w = 0.f;
			if (valid) {
				const float4 q0 = s_q0[buf][j];
				const float4 q1 = s_q1[buf][j];
				const float2 d = make_float2(q0.x - pixf.x, q0.y - pixf.y);
				const float power = -0.5f * (q0.z * d.x * d.x + q1.x * d.y * d.y) - q0.w * d.x * d.y;
				valid = !(power > 0.0f) && !(power < q1.z);  // q1.z: conservative bound below which alpha < 1/255 for sure
				float G = 0.f, alpha = 0.f;
				if (valid) {
					G = expf(power);
					alpha = fminf(0.99f, q1.y * G);
					valid = !(alpha < 1.0f / 255.0f);
				}
				if (valid) {
					const float4 q2 = s_q2[buf][j];
					T = T / (1.f - alpha);
					const float dchannel_dcolor = alpha * T;
					float dL_dopa = 0.0f;
					const float col[3] = {q2.x, q2.y, q2.z};
#pragma unroll
					for (int ch = 0; ch < 3; ch++) {
						const float c = col[ch];
						accum_rec[ch] = last_alpha * last_color[ch] + (1.f - last_alpha) * accum_rec[ch];
						last_color[ch] = c;
						const float dL_dchannel = dL_dpixel[ch];
						dL_dopa += (c - accum_rec[ch]) * dL_dchannel;
						v[7 + ch] = dchannel_dcolor * dL_dchannel;
					}
					if (SCH > 0) {
						const float *sp = semantics + (size_t)(s_id[buf][j] & idx_mask) * f.S;
#pragma unroll
						for (int ch = 0; ch < SCH; ch++)
							if (ch < f.S) {
								const float s = __ldg(sp + ch);
								accum_sem[ch] = last_alpha * last_sem[ch] + (1.f - last_alpha) * accum_sem[ch];
								last_sem[ch] = s;
								dL_dopa += (s - accum_sem[ch]) * dL_dsem_px[ch];
							}
						sem_w = dchannel_dcolor;
					}
					const float c_d = q1.w;
					accum_depth_rec = last_alpha * last_depth + (1.f - last_alpha) * accum_depth_rec;
					last_depth = c_d;
					dL_dopa += (c_d - accum_depth_rec) * dL_dpixel_depth;
					v[10] = dchannel_dcolor * dL_dpixel_depth;
					accum_alpha_rec = last_alpha + (1.f - last_alpha) * accum_alpha_rec;
					dL_dopa += (1 - accum_alpha_rec) * dL_dalpha_px;
					dL_dopa *= T;
					last_alpha = alpha;
					dL_dopa += (-T_final / (1.f - alpha)) * bg_dot_dpixel;
					const float dL_dG = q1.y * dL_dopa;
					const float gdx = G * d.x, gdy = G * d.y;
					const float dG_ddelx = -gdx * q0.z - gdy * q0.w;
					const float dG_ddely = -gdy * q1.x - gdx * q0.w;
					v[0] = dL_dG * dG_ddelx * ddelx_dx;
					v[1] = dL_dG * dG_ddely * ddely_dy;
					v[2] = fabsf(v[0]) + fabsf(v[1]);
					v[3] = -0.5f * gdx * d.x * dL_dG;
					v[4] = -0.5f * gdx * d.y * dL_dG;
					v[5] = -0.5f * gdy * d.y * dL_dG;
					v[6] = G * dL_dopa;
				}
			}
			const unsigned any = __ballot_sync(0xffffffffu, valid);
			if (any == 0u) continue;  // warp-uniform
			const float total = warp_reduce16_transposed(v, lane);
			const int comp = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
			if ((lane & 1) == 0 && comp < kNComp) s_part[warp][j][comp] = total;
			wmask |= 1ull << j;
			if (SCH > 0) {
				// feature channels: plain butterfly per channel, one global atomic per (warp, splat, channel)
				const uint32_t gid = s_id[buf][j] & idx_mask;
#pragma unroll
				for (int ch = 0; ch < SCH; ch++)
					if (ch < f.S) {
						float t = sem_w * dL_dsem_px[ch];
#pragma unroll
						for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
						if (lane == 0) atomicAdd(dL_dsemantics + (size_t)gid * f.S + ch, t);
					}
			}
		}
		if (lane == 0) s_mask[warp] = wmask;
		__syncthreads();
		// flush: thread -> (slot = tid/4, components 3*(tid%4) .. +2)
		{
			const int slot = tid >> 2, c0 = (tid & 3) * 3;
			if (slot < cnt) {
				float a0 = 0.f, a1 = 0.f, a2 = 0.f;
				bool live = false;
#pragma unroll
				for (int w = 0; w < 8; w++)
					if ((s_mask[w] >> slot) & 1ull) {
						live = true;
						a0 += s_part[w][slot][c0];
						a1 += s_part[w][slot][c0 + 1];
						a2 += s_part[w][slot][c0 + 2];
					}
				if (live) {
					float *dst = grad2d + (size_t)(s_id[buf][slot] & idx_mask) * kNComp + c0;
					atomicAdd(dst, a0);
					atomicAdd(dst + 1, a1);
					if (c0 + 2 < 11) atomicAdd(dst + 2, a2);
				}
			}
		}
		if (b + 1 < nb) stash((b + 1) & 1);
	}
}

cudaError_t launch_blend_bwd(const FrameDev &f, GeomView g, BinView b, ImgView img, const float *semantics,
                             const float *out_alpha, const float *dL_dcolor, const float *dL_ddepth, const float *dL_dalpha,
                             const float *dL_dsem, float *grad2d, float *dL_dsemantics, cudaStream_t st) {
	if (f.P == 0) return cudaSuccess;
	cudaError_t e = cudaMemsetAsync(grad2d, 0, (size_t)f.P * kNComp * sizeof(float), st);
	if (e != cudaSuccess) return e;
	if (f.S > 0 && (e = cudaMemsetAsync(dL_dsemantics, 0, (size_t)f.P * f.S * sizeof(float), st)) != cudaSuccess) return e;
	const int rows = band_rows(f.band);
	if (rows <= 0 || f.gx <= 0) return cudaSuccess;
	const dim3 grid(f.gx, rows);
#define SGR_LAUNCH_BWD(SCH)                                                                                                      \
	blend_bwd_kernel<SCH><<<grid, 256, 0, st>>>(f, img.ranges, b.vals_out, g.rec, semantics, img.n_contrib, img.tile_max_contrib, \
	                                            out_alpha, dL_dcolor, dL_ddepth, dL_dalpha, dL_dsem, grad2d, dL_dsemantics)
	if (f.S <= 0)
		SGR_LAUNCH_BWD(0);
	else if (f.S <= 4)
		SGR_LAUNCH_BWD(4);
	else if (f.S <= 8)
		SGR_LAUNCH_BWD(8);
	else if (f.S <= 16)
		SGR_LAUNCH_BWD(16);
	else
		SGR_LAUNCH_BWD(32);
#undef SGR_LAUNCH_BWD
	return cudaGetLastError();
}

}  // namespace sgr
