// losses.cu — the image-space training losses of street_gaussians, value AND gradient in two kernels.
//
// The reference builds  loss = (1 - l) * l1w * L1(image, gt, mask) + l * (1 - SSIM(image, gt, mask))  (train.py:101-104) out of
// lib/utils/loss_utils.py:21-37 (l1_loss) and :91-126 (ssim: five 11x11 Gaussian convolutions per image pair as grouped
// F.conv2d calls, plus ~15 elementwise kernels), then autograd replays all of it backwards to obtain dL/dimage — the tensor
// the rasterizer's backward consumes.  Here:
//   kernel 1 (ssim_stats_kernel): per 16x16 tile with a 5-pixel halo, separable 11-tap convolution of (x, y, x^2, y^2, xy) in
//            shared memory -> the SSIM map value, its three partial derivatives w.r.t. (mu_x, E[x^2], E[xy]) and the tile's
//            partial sums of SSIM, |x - y| over the mask and the mask count;
//   kernel 2 (ssim_grad_kernel): convolves the three derivative maps with the same window (the zero-padded Gaussian window
//            is symmetric, hence self-adjoint) and combines them with x, y, the L1 sign term and the mask into dL/dimage;
//            block (0,0,0) also finalises the scalars.
// Semantics follow the reference exactly: with a mask both images are ZEROED outside it before SSIM (loss_utils.py:95-97) and
// the SSIM mean runs over all pixels, while L1 averages over the masked pixels only (:31-35).
#include "sgr_common.cuh"

namespace sgr {

constexpr int kWin = 11, kHalo = 5, kTileL = 16, kExt = kTileL + 2 * kHalo;  // 26

struct GaussWin {
	float w[kWin];
};
// loss_utils.py:84-86: exp(-(x - 5)^2 / (2 * 1.5^2)) normalised (float32, like torch.Tensor([...]) / sum)
static GaussWin make_window() {
	GaussWin g;
	float s = 0.f;
	for (int k = 0; k < kWin; k++) {
		g.w[k] = (float)exp(-(double)((k - kWin / 2) * (k - kWin / 2)) / (2.0 * 1.5 * 1.5));
		s += g.w[k];
	}
	for (int k = 0; k < kWin; k++) g.w[k] /= s;
	return g;
}

// acc: [0] sum of the SSIM map, [1] sum |x - y| over masked elements, [2] number of masked PIXELS (counted on channel 0)
__global__ void __launch_bounds__(256) ssim_stats_kernel(const int C, const int H, const int W, const GaussWin win, const float *__restrict__ img,
                                                        const float *__restrict__ gt, const uint8_t *__restrict__ mask,
                                                        float *__restrict__ d_mu, float *__restrict__ d_xx, float *__restrict__ d_xy,
                                                        double *__restrict__ acc) {
	__shared__ float sx[kExt][kExt + 1], sy[kExt][kExt + 1];
	__shared__ float h[5][kExt][kTileL + 1];
	__shared__ float red[3][8];
	const int c = blockIdx.z, tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
	const int x0 = blockIdx.x * kTileL - kHalo, y0 = blockIdx.y * kTileL - kHalo;
	const size_t plane = (size_t)H * W;
	const float *ip = img + (size_t)c * plane, *gp = gt + (size_t)c * plane;
	for (int e = threadIdx.x; e < kExt * kExt; e += 256) {
		const int r = e / kExt, q = e - r * kExt, yy = y0 + r, xx = x0 + q;
		float a = 0.f, b = 0.f;
		if (yy >= 0 && yy < H && xx >= 0 && xx < W && (mask == nullptr || mask[(size_t)yy * W + xx])) {
			a = ip[(size_t)yy * W + xx];
			b = gp[(size_t)yy * W + xx];
		}
		sx[r][q] = a;
		sy[r][q] = b;
	}
	__syncthreads();
	// horizontal pass: 26 rows x 16 columns x 5 quantities
	for (int e = threadIdx.x; e < kExt * kTileL; e += 256) {
		const int r = e / kTileL, q = e - r * kTileL;
		float m1 = 0.f, m2 = 0.f, s11 = 0.f, s22 = 0.f, s12 = 0.f;
#pragma unroll
		for (int k = 0; k < kWin; k++) {
			const float a = sx[r][q + k], b = sy[r][q + k], w = win.w[k];
			m1 += w * a; m2 += w * b; s11 += w * a * a; s22 += w * b * b; s12 += w * a * b;
		}
		h[0][r][q] = m1; h[1][r][q] = m2; h[2][r][q] = s11; h[3][r][q] = s22; h[4][r][q] = s12;
	}
	__syncthreads();
	const int px = blockIdx.x * kTileL + tx, py = blockIdx.y * kTileL + ty;
	float v_ssim = 0.f, v_l1 = 0.f, v_cnt = 0.f;
	if (px < W && py < H) {
		float m1 = 0.f, m2 = 0.f, s11 = 0.f, s22 = 0.f, s12 = 0.f;
#pragma unroll
		for (int k = 0; k < kWin; k++) {
			const float w = win.w[k];
			m1 += w * h[0][ty + k][tx]; m2 += w * h[1][ty + k][tx]; s11 += w * h[2][ty + k][tx]; s22 += w * h[3][ty + k][tx];
			s12 += w * h[4][ty + k][tx];
		}
		const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
		const float mu1_sq = m1 * m1, mu2_sq = m2 * m2, mu12 = m1 * m2;
		const float sig1 = s11 - mu1_sq, sig2 = s22 - mu2_sq, sig12 = s12 - mu12;
		const float A1 = 2.f * mu12 + C1, A2 = 2.f * sig12 + C2, B1 = mu1_sq + mu2_sq + C1, B2 = sig1 + sig2 + C2;
		const float inv = 1.0f / (B1 * B2);
		const float S = A1 * A2 * inv;
		v_ssim = S;
		const size_t o = (size_t)c * plane + (size_t)py * W + px;
		// S = A1 A2 / (B1 B2);  dA1/dmu1 = 2 mu2, dA2/dmu1 = -2 mu2, dB1/dmu1 = 2 mu1, dB2/dmu1 = -2 mu1, dB2/ds11 = 1, dA2/ds12 = 2
		d_mu[o] = 2.f * m2 * (A2 - A1) * inv - S * 2.f * m1 * (1.0f / B1 - 1.0f / B2);
		d_xx[o] = -S / B2;
		d_xy[o] = 2.f * A1 * inv;
		const bool on = mask == nullptr || mask[(size_t)py * W + px];
		if (on) {
			v_l1 = fabsf(ip[(size_t)py * W + px] - gp[(size_t)py * W + px]);
			v_cnt = c == 0 ? 1.f : 0.f;
		}
	}
	// block reduction of the three partial sums
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) {
		v_ssim += __shfl_xor_sync(0xffffffffu, v_ssim, o);
		v_l1 += __shfl_xor_sync(0xffffffffu, v_l1, o);
		v_cnt += __shfl_xor_sync(0xffffffffu, v_cnt, o);
	}
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	if (lane == 0) { red[0][warp] = v_ssim; red[1][warp] = v_l1; red[2][warp] = v_cnt; }
	__syncthreads();
	if (threadIdx.x < 3) {
		double s = 0.0;
		for (int k = 0; k < 8; k++) s += (double)red[threadIdx.x][k];
		atomicAdd(acc + threadIdx.x, s);
	}
}

// dL/dimage = w_ssim/(C H W) * [ conv(d_mu) + 2 x conv(d_xx) + y conv(d_xy) ] + w_l1/(n_mask C) * sign(x - y), zero outside the mask.
// scalars: [0] w_l1 * L1 + w_ssim * SSIM, [1] L1, [2] SSIM, [3] masked pixel count
__global__ void __launch_bounds__(256) ssim_grad_kernel(const int C, const int H, const int W, const GaussWin win, const float *__restrict__ img,
                                                       const float *__restrict__ gt, const uint8_t *__restrict__ mask,
                                                       const float *__restrict__ d_mu, const float *__restrict__ d_xx,
                                                       const float *__restrict__ d_xy, const double *__restrict__ acc, const float w_l1,
                                                       const float w_ssim, float *__restrict__ grad, float *__restrict__ scalars) {
	__shared__ float s[3][kExt][kExt + 1];
	__shared__ float h[3][kExt][kTileL + 1];
	const int c = blockIdx.z, tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
	const int x0 = blockIdx.x * kTileL - kHalo, y0 = blockIdx.y * kTileL - kHalo;
	const size_t plane = (size_t)H * W;
	const double n_el = (double)C * (double)plane, n_l1 = acc[2] * (double)C;
	if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0 && scalars != nullptr) {
		const double l1 = n_l1 > 0.0 ? acc[1] / n_l1 : 0.0 / 0.0;  // (an empty mask is a NaN mean in the reference as well)
		const double ss = acc[0] / n_el;
		scalars[0] = (float)((double)w_l1 * l1 + (double)w_ssim * ss);
		scalars[1] = (float)l1;
		scalars[2] = (float)ss;
		scalars[3] = (float)acc[2];
	}
	if (grad == nullptr) return;
	const size_t cbase = (size_t)c * plane;
	for (int e = threadIdx.x; e < kExt * kExt; e += 256) {
		const int r = e / kExt, q = e - r * kExt, yy = y0 + r, xx = x0 + q;
		const bool in = yy >= 0 && yy < H && xx >= 0 && xx < W;
		const size_t o = cbase + (size_t)yy * W + xx;
		s[0][r][q] = in ? d_mu[o] : 0.f;
		s[1][r][q] = in ? d_xx[o] : 0.f;
		s[2][r][q] = in ? d_xy[o] : 0.f;
	}
	__syncthreads();
	for (int e = threadIdx.x; e < kExt * kTileL; e += 256) {
		const int r = e / kTileL, q = e - r * kTileL;
		float a = 0.f, b = 0.f, d = 0.f;
#pragma unroll
		for (int k = 0; k < kWin; k++) {
			const float w = win.w[k];
			a += w * s[0][r][q + k]; b += w * s[1][r][q + k]; d += w * s[2][r][q + k];
		}
		h[0][r][q] = a; h[1][r][q] = b; h[2][r][q] = d;
	}
	__syncthreads();
	const int px = blockIdx.x * kTileL + tx, py = blockIdx.y * kTileL + ty;
	if (px >= W || py >= H) return;
	const size_t o = cbase + (size_t)py * W + px;
	const bool on = mask == nullptr || mask[(size_t)py * W + px];
	float g = 0.f;
	if (on) {
		float a = 0.f, b = 0.f, d = 0.f;
#pragma unroll
		for (int k = 0; k < kWin; k++) {
			const float w = win.w[k];
			a += w * h[0][ty + k][tx]; b += w * h[1][ty + k][tx]; d += w * h[2][ty + k][tx];
		}
		const float x = img[o], y = gt[o];
		g = (float)((double)w_ssim / n_el) * (a + 2.f * x * b + y * d);
		const float df = x - y;
		const float sgn = df > 0.f ? 1.f : (df < 0.f ? -1.f : 0.f);  // torch.abs backward: sign(0) = 0
		if (n_l1 > 0.0) g += (float)((double)w_l1 / n_l1) * sgn;
	}
	grad[o] = g;
}

// sky / accumulation loss (train.py:107-113): acc clamped to [1e-6, 1 - 1e-6]; mean of  sky ? -log(1 - acc) : -log(acc).
// out[0] += sum; grad = weight / N * d/dacc (zero where the clamp is active, like torch.clamp's backward).
__global__ void __launch_bounds__(256) sky_loss_kernel(const size_t N, const float *__restrict__ accm, const uint8_t *__restrict__ sky, const float weight,
                                                      float *__restrict__ grad, double *__restrict__ out) {
	__shared__ float red[8];
	float v = 0.f;
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (size_t)gridDim.x * blockDim.x) {
		const float a = accm[i];
		const float ac = fminf(fmaxf(a, 1e-6f), 1.f - 1e-6f);
		const bool inside = a >= 1e-6f && a <= 1.f - 1e-6f;
		const bool s = sky[i] != 0;
		v += s ? -logf(1.f - ac) : -logf(ac);
		if (grad) grad[i] = inside ? (weight / (float)N) * (s ? 1.f / (1.f - ac) : -1.f / ac) : 0.f;
	}
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
	__syncthreads();
	if (threadIdx.x == 0) {
		double s = 0.0;
		for (int k = 0; k < 8; k++) s += (double)red[k];
		atomicAdd(out, s);
	}
}
__global__ void sky_finalize_kernel(const size_t N, const float weight, const double *__restrict__ sum, float *__restrict__ scalars) {
	scalars[0] = (float)((double)weight * (*sum / (double)N));
	scalars[1] = (float)(*sum / (double)N);
}

size_t image_loss_scratch_bytes(int C, int H, int W) { return align_up((size_t)3 * C * H * W * sizeof(float)) + 256; }

cudaError_t launch_image_loss(int C, int H, int W, const float *img, const float *gt, const uint8_t *mask, float w_l1, float w_ssim, float *grad,
                              float *scalars, void *scratch, cudaStream_t st) {
	static const GaussWin win = make_window();
	const size_t n = (size_t)C * H * W;
	float *d_mu = reinterpret_cast<float *>(scratch), *d_xx = d_mu + n, *d_xy = d_xx + n;
	double *acc = reinterpret_cast<double *>(reinterpret_cast<char *>(scratch) + align_up(3 * n * sizeof(float)));
	cudaError_t e = cudaMemsetAsync(acc, 0, 4 * sizeof(double), st);
	if (e != cudaSuccess) return e;
	const dim3 grid((W + kTileL - 1) / kTileL, (H + kTileL - 1) / kTileL, C);
	count_launch(2);
	ssim_stats_kernel<<<grid, 256, 0, st>>>(C, H, W, win, img, gt, mask, d_mu, d_xx, d_xy, acc);
	ssim_grad_kernel<<<grad ? grid : dim3(1, 1, 1), 256, 0, st>>>(C, H, W, win, img, gt, mask, d_mu, d_xx, d_xy, acc, w_l1, w_ssim, grad, scalars);
	return cudaGetLastError();
}

cudaError_t launch_sky_loss(size_t N, const float *accm, const uint8_t *sky, float weight, float *grad, float *scalars, void *scratch, cudaStream_t st) {
	double *sum = reinterpret_cast<double *>(scratch);
	cudaError_t e = cudaMemsetAsync(sum, 0, sizeof(double), st);
	if (e != cudaSuccess) return e;
	const unsigned nblk = (unsigned)((N + 255) / 256 < 148 * 8 ? (N + 255) / 256 : 148 * 8);
	count_launch(2);
	sky_loss_kernel<<<nblk ? nblk : 1, 256, 0, st>>>(N, accm, sky, weight, grad, sum);
	sky_finalize_kernel<<<1, 1, 0, st>>>(N, weight, sum, scalars);
	return cudaGetLastError();
}

}  // namespace sgr
