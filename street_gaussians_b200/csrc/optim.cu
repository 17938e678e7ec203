// optim.cu — what runs right after the rasterizer's backward in every training iteration:
//
//  * densification statistics (lib/models/street_gaussian_model.py:551-571): per sub-model, for the Gaussians that were visible
//    (radii > 0):  max_radii2D = max(max_radii2D, radii),  xyz_gradient_accum[:,0] += |grad2D.xy|,  [:,1] += |grad2D.z|,  denom += 1.
//    The reference slices the composed arrays per model and runs ~8 indexed PyTorch kernels per model per iteration; here ONE
//    kernel walks the composed index space with the same segment table the composer uses.
//  * the optimiser step (lib/models/gaussian_model.py:316-318 -> torch.optim.Adam(lr per group, eps = 1e-15), :300-303): one
//    multi-tensor Adam kernel over every parameter tensor of every sub-model (the reference: 6 groups x (1 + #actors) models,
//    each a handful of foreach kernels), same arithmetic as torch.optim.Adam without weight decay / amsgrad:
//        m <- m + (g - m)(1 - b1);  v <- b2 v + (1 - b2) g^2;  p <- p - (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
#include "sgr_common.cuh"

namespace sgr {

constexpr int kStatSeg = SGR_MAX_SEGMENTS_PER_LAUNCH;
struct StatTable {
	int n;
	int start[kStatSeg + 1];
	float *max_radii2D[kStatSeg];
	float *grad_accum[kStatSeg];  // [count, 2]
	float *denom[kStatSeg];       // [count, 1]
};

__global__ void __launch_bounds__(256) densify_stats_kernel(const StatTable t, const int32_t *__restrict__ radii,
                                                           const float *__restrict__ grad2d /* viewspace_points.grad [P,3] */) {
	const int i = t.start[0] + blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= t.start[t.n]) return;
	const int r = radii[i];
	if (r <= 0) return;  // visibility_filter = radii > 0 (street_gaussian_renderer.py:274)
	int lo = 0, hi = t.n - 1;
	while (lo < hi) {
		const int mid = (lo + hi + 1) >> 1;
		if (t.start[mid] <= i) lo = mid; else hi = mid - 1;
	}
	const size_t l = (size_t)(i - t.start[lo]);
	float *mr = t.max_radii2D[lo] + l;
	*mr = fmaxf(*mr, (float)r);
	const float gx = grad2d[3 * (size_t)i], gy = grad2d[3 * (size_t)i + 1], gz = grad2d[3 * (size_t)i + 2];
	t.grad_accum[lo][2 * l] += sqrtf(gx * gx + gy * gy);  // torch.norm(grad[:, :2], dim=-1)
	t.grad_accum[lo][2 * l + 1] += fabsf(gz);             // torch.norm(grad[:, 2:], dim=-1) of one column
	t.denom[lo][l] += 1.f;
}

cudaError_t launch_densify_stats(const SgrStatSegment *segs, int nseg, const int32_t *radii, const float *grad2d, cudaStream_t st) {
	for (int first = 0; first < nseg; first += kStatSeg) {
		StatTable t;
		const int n = nseg - first < kStatSeg ? nseg - first : kStatSeg;
		t.n = n;
		for (int k = 0; k < n; k++) {
			const SgrStatSegment &s = segs[first + k];
			t.start[k] = s.start;
			t.max_radii2D[k] = s.max_radii2D; t.grad_accum[k] = s.xyz_gradient_accum; t.denom[k] = s.denom;
		}
		t.start[n] = segs[first + n - 1].start + segs[first + n - 1].count;
		const int count = t.start[n] - t.start[0];
		if (count <= 0) continue;
		count_launch();
		densify_stats_kernel<<<(count + 255) / 256, 256, 0, st>>>(t, radii, grad2d);
	}
	return cudaGetLastError();
}

// ---- multi-tensor Adam ----
constexpr int kAdamTensors = 48;         // tensors per launch (kernel-parameter table)
constexpr int kAdamChunk = 256 * 4 * 8;  // elements per block
struct AdamTable {
	int n;
	float *p[kAdamTensors];
	const float *g[kAdamTensors];
	float *m[kAdamTensors], *v[kAdamTensors];
	long long numel[kAdamTensors];
	float step_size[kAdamTensors];   // lr / (1 - b1^t)
	float inv_sqrt_bc2[kAdamTensors];  // 1 / sqrt(1 - b2^t)
	int block_start[kAdamTensors + 1];  // first block of each tensor
};

// omb1 = fl(1 - beta1), omb2 = fl(1 - beta2) formed in double on the host, exactly the scalars torch hands to lerp_ / addcmul_
__global__ void __launch_bounds__(256) adam_kernel(const AdamTable t, const float omb1, const float b2, const float omb2, const float eps) {
	int k = 0;  // tensor of this block (linear search: <= 48 entries, warp-uniform)
	while (k + 1 < t.n && (int)blockIdx.x >= t.block_start[k + 1]) k++;
	const long long base = (long long)(blockIdx.x - t.block_start[k]) * kAdamChunk;
	const long long n = t.numel[k];
	float *__restrict__ p = t.p[k];
	const float *__restrict__ g = t.g[k];
	float *__restrict__ m = t.m[k], *__restrict__ v = t.v[k];
	const float ss = t.step_size[k], ibc2 = t.inv_sqrt_bc2[k];
	for (long long i = base + threadIdx.x; i < base + kAdamChunk && i < n; i += 256) {
		const float gi = g[i];
		const float mi = m[i] + (gi - m[i]) * omb1;              // exp_avg.lerp_(grad, 1 - beta1)
		const float vi = v[i] * b2 + omb2 * gi * gi;             // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value = 1 - beta2)
		m[i] = mi;
		v[i] = vi;
		p[i] = p[i] - ss * (mi / (sqrtf(vi) * ibc2 + eps));      // param.addcdiv_(exp_avg, sqrt(v)/sqrt(bc2) + eps, value = -step_size)
	}
}

cudaError_t launch_adam(const SgrAdamTensor *ts, int n_tensors, double beta1, double beta2, double eps, cudaStream_t st) {
	for (int first = 0; first < n_tensors; first += kAdamTensors) {
		AdamTable t;
		const int n = n_tensors - first < kAdamTensors ? n_tensors - first : kAdamTensors;
		t.n = n;
		int blocks = 0;
		for (int k = 0; k < n; k++) {
			const SgrAdamTensor &a = ts[first + k];
			t.p[k] = a.param; t.g[k] = a.grad; t.m[k] = a.exp_avg; t.v[k] = a.exp_avg_sq; t.numel[k] = a.numel;
			const double bc1 = 1.0 - pow(beta1, (double)a.step), bc2 = 1.0 - pow(beta2, (double)a.step);
			t.step_size[k] = (float)((double)a.lr / bc1);
			t.inv_sqrt_bc2[k] = (float)(1.0 / sqrt(bc2));
			t.block_start[k] = blocks;
			blocks += (int)((a.numel + kAdamChunk - 1) / kAdamChunk);
		}
		t.block_start[n] = blocks;
		if (blocks == 0) continue;
		count_launch();
		adam_kernel<<<blocks, 256, 0, st>>>(t, (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2), (float)eps);
	}
	return cudaGetLastError();
}

}  // namespace sgr
