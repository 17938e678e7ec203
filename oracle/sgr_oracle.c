/*
 * sgr_oracle.c — CPU restatement of the reference differentiable Gaussian rasterizer.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under street_gaussians_b200/ may import, link or call this
 * file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
 *
 * It restates, in plain C (fp32 per-pair arithmetic, fp64 only where the reference itself
 * promotes to double, and fp64 for the per-Gaussian gradient SUMS so the result does not depend
 * on summation order), the algorithm of /root/reference/submodules/diff-gaussian-rasterization
 * (abbreviated DGR/ below).  Every function cites the reference file:line it follows.
 *
 * Pinning: the reference ships no golden vectors.  This oracle is pinned
 * against outputs of the reference itself (oracle/_ref, built from the unmodified sources by
 * oracle/build_ref.sh and run on a B200) stored as fixtures under tests/golden/ by
 * tests/golden/make_golden.py.
 *
 * Build: gcc -O2 -fopenmp -shared -fPIC -ffp-contract=off sgr_oracle.c -o _build/libsgr_oracle.so -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define TILE 16 /* DGR/cuda_rasterizer/config.h:17-18 BLOCK_X = BLOCK_Y = 16 */

typedef struct {
	int32_t P, D, M, S, W, H;
	float tanfovx, tanfovy, scale_modifier;
	float bg[3];
	float view[16]; /* W2C^T row-major == column-major W2C, as the kernels index it */
	float proj[16];
	float campos[3];
} OrParams;

typedef struct {
	OrParams prm;
	int gx, gy;
	/* per-Gaussian forward state (the reference's GeometryState, rasterizer_impl.h:21-37) */
	float *depth, *xy, *conic_op, *rgb, *cov3d;
	uint8_t *clamped;
	int32_t *radii;
	uint32_t *tiles;
	/* binning */
	int64_t R;
	uint32_t *point_list;
	int64_t *range_lo, *range_hi;
	/* per-pixel */
	uint32_t *n_contrib;
	/* evaluated-pair statistics (for flop accounting) */
	int64_t pairs_evaluated, pairs_blended;
} OrState;

/* SH basis constants: DGR/cuda_rasterizer/auxiliary.h:22-39 */
static const float C0 = 0.28209479177387814f, C1 = 0.4886025119029199f;
static const float C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                            -1.0925484305920792f, 0.5462742152960396f};
static const float C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f, 0.3731763325901154f,
                            -0.4570457994644658f, 1.445305721320277f, -0.5900435899266435f};

/* auxiliary.h:58-77 — the matrices are indexed column-major */
static void xform4x3(const float *p, const float *m, float *o) {
	o[0] = m[0] * p[0] + m[4] * p[1] + m[8] * p[2] + m[12];
	o[1] = m[1] * p[0] + m[5] * p[1] + m[9] * p[2] + m[13];
	o[2] = m[2] * p[0] + m[6] * p[1] + m[10] * p[2] + m[14];
}
static void xform4x4(const float *p, const float *m, float *o) {
	xform4x3(p, m, o);
	o[3] = m[3] * p[0] + m[7] * p[1] + m[11] * p[2] + m[15];
}

/* auxiliary.h:41-44 — NB the literals 1.0 and 0.5 are doubles, so the reference evaluates in fp64 */
static float ndc2pix(float v, int S) { return (float)(((v + 1.0) * S - 1.0) * 0.5); }

static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* auxiliary.h:46-56 — (int) truncation toward zero BEFORE the clamp */
static void get_rect(const float *p, int r, int gx, int gy, int *mn, int *mx) {
	mn[0] = clampi((int)((p[0] - r) / TILE), 0, gx);
	mn[1] = clampi((int)((p[1] - r) / TILE), 0, gy);
	mx[0] = clampi((int)((p[0] + r + TILE - 1) / TILE), 0, gx);
	mx[1] = clampi((int)((p[1] + r + TILE - 1) / TILE), 0, gy);
}

/* forward.cu:118-152 — quaternion is NOT normalised (line 127) */
static void rot_from_quat(const float *q, float Rm[3][3]) {
	float r = q[0], x = q[1], y = q[2], z = q[3];
	Rm[0][0] = 1.f - 2.f * (y * y + z * z); Rm[0][1] = 2.f * (x * y - r * z); Rm[0][2] = 2.f * (x * z + r * y);
	Rm[1][0] = 2.f * (x * y + r * z); Rm[1][1] = 1.f - 2.f * (x * x + z * z); Rm[1][2] = 2.f * (y * z - r * x);
	Rm[2][0] = 2.f * (x * z - r * y); Rm[2][1] = 2.f * (y * z + r * x); Rm[2][2] = 1.f - 2.f * (x * x + y * y);
}
static void cov3d_from_scale_rot(const float *s, float mod, const float *q, float *c6) {
	float Rm[3][3], Mk[3][3];
	rot_from_quat(q, Rm);
	/* M = S * R_glm, M_{k,i} = (mod*s_k) * Rstd_{i,k}; Sigma = M^T M */
	for (int k = 0; k < 3; k++)
		for (int i = 0; i < 3; i++) Mk[k][i] = (mod * s[k]) * Rm[i][k];
	float Sg[3][3];
	for (int i = 0; i < 3; i++)
		for (int j = 0; j < 3; j++) Sg[i][j] = Mk[0][i] * Mk[0][j] + Mk[1][i] * Mk[1][j] + Mk[2][i] * Mk[2][j];
	c6[0] = Sg[0][0]; c6[1] = Sg[0][1]; c6[2] = Sg[0][2]; c6[3] = Sg[1][1]; c6[4] = Sg[1][2]; c6[5] = Sg[2][2];
}

/* Shared by forward (forward.cu:74-113) and backward (backward.cu:144-200): A = J*Rw (2x3),
 * clamped t, gating flags.  A[i][r] is the reference's T[r][i]. */
typedef struct { float A[2][3]; float t[3]; float gx, gy; } Proj2;
static void proj_jacobian(const float *mean, const OrParams *pm, float fx, float fy, Proj2 *o) {
	float t[3];
	xform4x3(mean, pm->view, t);
	const float limx = 1.3f * pm->tanfovx, limy = 1.3f * pm->tanfovy;
	const float txtz = t[0] / t[2], tytz = t[1] / t[2];
	t[0] = fminf(limx, fmaxf(-limx, txtz)) * t[2];
	t[1] = fminf(limy, fmaxf(-limy, tytz)) * t[2];
	o->gx = (txtz < -limx || txtz > limx) ? 0.f : 1.f;
	o->gy = (tytz < -limy || tytz > limy) ? 0.f : 1.f;
	const float J00 = fx / t[2], J02 = -(fx * t[0]) / (t[2] * t[2]);
	const float J11 = fy / t[2], J12 = -(fy * t[1]) / (t[2] * t[2]);
	const float *v = pm->view;
	for (int r = 0; r < 3; r++) {
		o->A[0][r] = v[4 * r + 0] * J00 + v[4 * r + 1] * 0.0f + v[4 * r + 2] * J02;
		o->A[1][r] = v[4 * r + 0] * 0.0f + v[4 * r + 1] * J11 + v[4 * r + 2] * J12;
	}
	o->t[0] = t[0]; o->t[1] = t[1]; o->t[2] = t[2];
}
static void cov2d_from(const Proj2 *pj, const float *c6, float *abc) {
	const float V[3][3] = {{c6[0], c6[1], c6[2]}, {c6[1], c6[3], c6[4]}, {c6[2], c6[4], c6[5]}};
	/* B = T^T Vrk^T (rows 0,1 only), cov = B T */
	float B[2][3];
	for (int i = 0; i < 2; i++)
		for (int j = 0; j < 3; j++) B[i][j] = pj->A[i][0] * V[j][0] + pj->A[i][1] * V[j][1] + pj->A[i][2] * V[j][2];
	float c00 = B[0][0] * pj->A[0][0] + B[0][1] * pj->A[0][1] + B[0][2] * pj->A[0][2];
	float c10 = B[1][0] * pj->A[0][0] + B[1][1] * pj->A[0][1] + B[1][2] * pj->A[0][2];
	float c11 = B[1][0] * pj->A[1][0] + B[1][1] * pj->A[1][1] + B[1][2] * pj->A[1][2];
	abc[0] = c00 + 0.3f; abc[1] = c10; abc[2] = c11 + 0.3f; /* forward.cu:110-112 low-pass */
}

/* forward.cu:20-71 */
static void sh_to_rgb(int deg, const float *pos, const float *campos, const float *sh, float *rgb, uint8_t *clamped) {
	float d[3] = {pos[0] - campos[0], pos[1] - campos[1], pos[2] - campos[2]};
	float len = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
	float x = d[0] / len, y = d[1] / len, z = d[2] / len;
	for (int c = 0; c < 3; c++) {
#define SHC(k) sh[(k) * 3 + c]
		float res = C0 * SHC(0);
		if (deg > 0) {
			res = res - C1 * y * SHC(1) + C1 * z * SHC(2) - C1 * x * SHC(3);
			if (deg > 1) {
				float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
				res = res + C2[0] * xy * SHC(4) + C2[1] * yz * SHC(5) + C2[2] * (2.0f * zz - xx - yy) * SHC(6) +
				      C2[3] * xz * SHC(7) + C2[4] * (xx - yy) * SHC(8);
				if (deg > 2) {
					res = res + C3[0] * y * (3.0f * xx - yy) * SHC(9) + C3[1] * xy * z * SHC(10) +
					      C3[2] * y * (4.0f * zz - xx - yy) * SHC(11) +
					      C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * SHC(12) +
					      C3[4] * x * (4.0f * zz - xx - yy) * SHC(13) + C3[5] * z * (xx - yy) * SHC(14) +
					      C3[6] * x * (xx - 3.0f * yy) * SHC(15);
				}
			}
		}
#undef SHC
		res += 0.5f;
		clamped[c] = res < 0;
		rgb[c] = fmaxf(res, 0.0f);
	}
}

void or_free(OrState *st) {
	if (!st) return;
	free(st->depth); free(st->xy); free(st->conic_op); free(st->rgb); free(st->cov3d); free(st->clamped);
	free(st->radii); free(st->tiles); free(st->point_list); free(st->range_lo); free(st->range_hi); free(st->n_contrib);
	free(st);
}

/* Per-Gaussian forward: forward.cu:155-256 (+ in_frustum auxiliary.h:139-164). */
static void preprocess_all(OrState *st, const float *means3D, const float *shs, const float *colors_precomp,
                           const float *opacities, const float *scales, const float *rotations,
                           const float *cov3D_precomp) {
	const OrParams *pm = &st->prm;
	const int P = pm->P;
	const float fy = pm->H / (2.0f * pm->tanfovy), fx = pm->W / (2.0f * pm->tanfovx); /* rasterizer_impl.cu:225-226 */
#pragma omp parallel for schedule(static)
	for (int i = 0; i < P; i++) {
		st->radii[i] = 0; st->tiles[i] = 0;
		const float *p = means3D + 3 * i;
		float ph[4], pv[3];
		xform4x4(p, pm->proj, ph);
		float pw = 1.0f / (ph[3] + 0.0000001f);
		float pproj[2] = {ph[0] * pw, ph[1] * pw};
		xform4x3(p, pm->view, pv);
		if (pv[2] <= 0.2f) continue; /* auxiliary.h:154 near cull only */
		float *c6 = st->cov3d + 6 * i;
		if (cov3D_precomp) memcpy(c6, cov3D_precomp + 6 * i, 6 * sizeof(float));
		else cov3d_from_scale_rot(scales + 3 * i, pm->scale_modifier, rotations + 4 * i, c6);
		Proj2 pj; float abc[3];
		proj_jacobian(p, pm, fx, fy, &pj);
		cov2d_from(&pj, c6, abc);
		float det = abc[0] * abc[2] - abc[1] * abc[1];
		if (det == 0.0f) continue;
		float det_inv = 1.f / det;
		float conic[3] = {abc[2] * det_inv, -abc[1] * det_inv, abc[0] * det_inv};
		float mid = 0.5f * (abc[0] + abc[2]);
		float l1 = mid + sqrtf(fmaxf(0.1f, mid * mid - det));
		float l2 = mid - sqrtf(fmaxf(0.1f, mid * mid - det));
		float my_radius = ceilf(3.f * sqrtf(fmaxf(l1, l2)));
		float pix[2] = {ndc2pix(pproj[0], pm->W), ndc2pix(pproj[1], pm->H)};
		int mn[2], mx[2];
		get_rect(pix, (int)my_radius, st->gx, st->gy, mn, mx);
		if ((mx[0] - mn[0]) * (mx[1] - mn[1]) == 0) continue;
		if (colors_precomp) memcpy(st->rgb + 3 * i, colors_precomp + 3 * i, 3 * sizeof(float));
		else sh_to_rgb(pm->D, p, pm->campos, shs + (size_t)i * pm->M * 3, st->rgb + 3 * i, st->clamped + 3 * i);
		st->depth[i] = pv[2];
		st->radii[i] = (int)my_radius;
		st->xy[2 * i] = pix[0]; st->xy[2 * i + 1] = pix[1];
		st->conic_op[4 * i] = conic[0]; st->conic_op[4 * i + 1] = conic[1]; st->conic_op[4 * i + 2] = conic[2];
		st->conic_op[4 * i + 3] = opacities[i];
		st->tiles[i] = (uint32_t)((mx[1] - mn[1]) * (mx[0] - mn[0]));
	}
}

typedef struct { uint64_t key; uint32_t val; } KV;

/* LSD radix sort, 8 bits per pass, stable — same ordering semantics as cub::DeviceRadixSort::SortPairs
 * (rasterizer_impl.cu:303-311): ascending key, equal keys keep emission order. */
static void radix_sort_kv(KV *a, KV *tmp, int64_t n, int bits) {
	for (int sh = 0; sh < bits; sh += 8) {
		int64_t cnt[257] = {0};
		for (int64_t i = 0; i < n; i++) cnt[((a[i].key >> sh) & 0xFF) + 1]++;
		for (int b = 0; b < 256; b++) cnt[b + 1] += cnt[b];
		for (int64_t i = 0; i < n; i++) tmp[cnt[(a[i].key >> sh) & 0xFF]++] = a[i];
		KV *t = a; a = tmp; tmp = t;
	}
	if ((bits + 7) / 8 % 2) memcpy(tmp, a, n * sizeof(KV)); /* result must end in the caller's `a` */
}

/* rasterizer_impl.cu:70-138, 278-321 */
static int bin_all(OrState *st) {
	const int P = st->prm.P;
	int64_t R = 0;
	for (int i = 0; i < P; i++) R += st->tiles[i];
	st->R = R;
	KV *kv = (KV *)malloc((R + 1) * sizeof(KV)), *tmp = (KV *)malloc((R + 1) * sizeof(KV));
	if (!kv || !tmp) return -1;
	int64_t off = 0;
	for (int i = 0; i < P; i++) {
		if (st->radii[i] <= 0) continue;
		int mn[2], mx[2];
		get_rect(st->xy + 2 * i, st->radii[i], st->gx, st->gy, mn, mx);
		uint32_t dbits; memcpy(&dbits, st->depth + i, 4);
		for (int y = mn[1]; y < mx[1]; y++)
			for (int x = mn[0]; x < mx[0]; x++) {
				kv[off].key = ((uint64_t)(y * st->gx + x) << 32) | dbits;
				kv[off].val = (uint32_t)i; off++;
			}
	}
	KV *orig = kv;
	radix_sort_kv(kv, tmp, R, 64);
	(void)orig;
	st->point_list = (uint32_t *)malloc((R + 1) * sizeof(uint32_t));
	const int nt = st->gx * st->gy;
	st->range_lo = (int64_t *)calloc(nt, sizeof(int64_t));
	st->range_hi = (int64_t *)calloc(nt, sizeof(int64_t));
	for (int64_t i = 0; i < R; i++) {
		st->point_list[i] = kv[i].val;
		uint32_t tcur = (uint32_t)(kv[i].key >> 32);
		if (i == 0) st->range_lo[tcur] = 0;
		else {
			uint32_t tprev = (uint32_t)(kv[i - 1].key >> 32);
			if (tcur != tprev) { st->range_hi[tprev] = i; st->range_lo[tcur] = i; }
		}
		if (i == R - 1) st->range_hi[tcur] = R;
	}
	free(kv); free(tmp);
	return 0;
}

/* Per-pixel front-to-back blend: forward.cu:340-467. */
static void blend_all(OrState *st, const float *semantics, float *out_color, float *out_depth, float *out_alpha,
                      float *out_sem) {
	const OrParams *pm = &st->prm;
	const int W = pm->W, H = pm->H, S = pm->S;
	const size_t HW = (size_t)W * H;
	int64_t ev = 0, bl = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : ev, bl)
	for (int tile = 0; tile < st->gx * st->gy; tile++) {
		const int tx = tile % st->gx, ty = tile / st->gx;
		const int64_t lo = st->range_lo[tile], hi = st->range_hi[tile];
		for (int py = ty * TILE; py < ty * TILE + TILE && py < H; py++)
			for (int px = tx * TILE; px < tx * TILE + TILE && px < W; px++) {
				const size_t pid = (size_t)W * py + px;
				const float pxf = (float)px, pyf = (float)py;
				float T = 1.0f, C[3] = {0, 0, 0}, weight = 0, Dp = 0;
				uint32_t contributor = 0, last = 0;
				for (int ch = 0; ch < S; ch++) out_sem[ch * HW + pid] = 0.f;
				for (int64_t k = lo; k < hi; k++) {
					contributor++; ev++;
					const uint32_t g = st->point_list[k];
					const float dx = st->xy[2 * g] - pxf, dy = st->xy[2 * g + 1] - pyf;
					const float *co = st->conic_op + 4 * g;
					const float power = -0.5f * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
					if (power > 0.0f) continue;
					const float alpha = fminf(0.99f, co[3] * expf(power));
					if (alpha < 1.0f / 255.0f) continue;
					const float test_T = T * (1 - alpha);
					if (test_T < 0.0001f) break; /* done=true: nothing later is visited (forward.cu:432-436) */
					for (int ch = 0; ch < 3; ch++) C[ch] += st->rgb[3 * g + ch] * alpha * T;
					for (int ch = 0; ch < S; ch++) out_sem[ch * HW + pid] += semantics[(size_t)g * S + ch] * alpha * T;
					weight += alpha * T;
					Dp += st->depth[g] * alpha * T;
					T = test_T;
					last = contributor; bl++;
				}
				st->n_contrib[pid] = last;
				for (int ch = 0; ch < 3; ch++) out_color[ch * HW + pid] = C[ch] + T * pm->bg[ch];
				out_alpha[pid] = weight;
				out_depth[pid] = Dp;
			}
	}
	st->pairs_evaluated = ev; st->pairs_blended = bl;
}

/* Whole forward: rasterizer_impl.cu:197-343.  Pointers may be NULL exactly where the reference accepts
 * empty tensors (shs xor colors_precomp; (scales,rotations) xor cov3D_precomp; semantics when S==0). */
OrState *or_forward(const OrParams *pm, const float *means3D, const float *shs, const float *colors_precomp,
                    const float *semantics, const float *opacities, const float *scales, const float *rotations,
                    const float *cov3D_precomp, float *out_color, float *out_depth, float *out_alpha, float *out_sem,
                    int32_t *radii_out) {
	OrState *st = (OrState *)calloc(1, sizeof(OrState));
	st->prm = *pm;
	const int P = pm->P;
	st->gx = (pm->W + TILE - 1) / TILE; st->gy = (pm->H + TILE - 1) / TILE;
	const size_t n = (size_t)(P > 0 ? P : 1);
	st->depth = (float *)calloc(n, 4); st->xy = (float *)calloc(2 * n, 4); st->conic_op = (float *)calloc(4 * n, 4);
	st->rgb = (float *)calloc(3 * n, 4); st->cov3d = (float *)calloc(6 * n, 4); st->clamped = (uint8_t *)calloc(3 * n, 1);
	st->radii = (int32_t *)calloc(n, 4); st->tiles = (uint32_t *)calloc(n, 4);
	st->n_contrib = (uint32_t *)calloc((size_t)pm->W * pm->H + 1, 4);
	preprocess_all(st, means3D, shs, colors_precomp, opacities, scales, rotations, cov3D_precomp);
	if (bin_all(st)) { or_free(st); return NULL; }
	blend_all(st, semantics, out_color, out_depth, out_alpha, out_sem);
	if (radii_out) memcpy(radii_out, st->radii, (size_t)P * 4);
	return st;
}

/* accessors used by the geometry-level parity tests */
int64_t or_num_rendered(const OrState *st) { return st->R; }
int64_t or_pairs_evaluated(const OrState *st) { return st->pairs_evaluated; }
int64_t or_pairs_blended(const OrState *st) { return st->pairs_blended; }
void or_get_geom(const OrState *st, float *depth, float *xy, float *conic_op, float *rgb, uint8_t *clamped,
                 uint32_t *tiles, float *cov3d) {
	const size_t P = st->prm.P;
	if (depth) memcpy(depth, st->depth, P * 4);
	if (xy) memcpy(xy, st->xy, 2 * P * 4);
	if (conic_op) memcpy(conic_op, st->conic_op, 4 * P * 4);
	if (rgb) memcpy(rgb, st->rgb, 3 * P * 4);
	if (clamped) memcpy(clamped, st->clamped, 3 * P);
	if (tiles) memcpy(tiles, st->tiles, P * 4);
	if (cov3d) memcpy(cov3d, st->cov3d, 6 * P * 4);
}
void or_get_n_contrib(const OrState *st, uint32_t *out) { memcpy(out, st->n_contrib, (size_t)st->prm.W * st->prm.H * 4); }

static inline void datomic(double *p, double v) {
#pragma omp atomic
	*p += v;
}

/* Backward.  Per-pixel part: backward.cu:415-641; per-Gaussian part: backward.cu:144-274, 278-341, 346-412, 20-139.
 * acc = P x (11+S) doubles: [0..2]=mean2D (x, y, |x|+|y|), [3..5]=conic (x, y, w), 6=opacity, [7..9]=color, 10=depth, 11..=sem */
int or_backward(const OrState *st, const float *means3D, const float *shs, const float *colors_precomp,
                const float *semantics, const float *scales, const float *rotations, const float *cov3D_precomp,
                const float *alphas, const float *dL_dcolor, const float *dL_ddepth_px, const float *dL_dalpha_px,
                const float *dL_dsem_px, float *g_means3D, float *g_means2D, float *g_sh, float *g_colors,
                float *g_sem, float *g_opacity, float *g_scales, float *g_rot, float *g_cov3D) {
	const OrParams *pm = &st->prm;
	const int P = pm->P, W = pm->W, H = pm->H, S = pm->S, M = pm->M;
	const size_t HW = (size_t)W * H;
	const int NA = 11 + S;
	double *acc = (double *)calloc((size_t)(P > 0 ? P : 1) * NA, sizeof(double));
	if (!acc) return -1;
	const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H; /* backward.cu:501-502 */
	const float *colors = colors_precomp ? colors_precomp : st->rgb;

#pragma omp parallel for schedule(dynamic, 1)
	for (int tile = 0; tile < st->gx * st->gy; tile++) {
		const int tx = tile % st->gx, ty = tile / st->gx;
		const int64_t lo = st->range_lo[tile];
		float *accum_sem = (float *)malloc(sizeof(float) * (2 * S + 1)), *last_sem = accum_sem + S;
		for (int py = ty * TILE; py < ty * TILE + TILE && py < H; py++)
			for (int px = tx * TILE; px < tx * TILE + TILE && px < W; px++) {
				const size_t pid = (size_t)W * py + px;
				const float pxf = (float)px, pyf = (float)py;
				const float T_final = 1 - alphas[pid];
				float T = T_final;
				const int64_t last_contributor = st->n_contrib[pid];
				float accum_rec[3] = {0, 0, 0}, last_color[3] = {0, 0, 0};
				float accum_depth = 0, last_depth = 0, accum_alpha = 0, last_alpha = 0;
				for (int ch = 0; ch < S; ch++) accum_sem[ch] = last_sem[ch] = 0;
				float dpix[3] = {dL_dcolor[pid], dL_dcolor[HW + pid], dL_dcolor[2 * HW + pid]};
				const float dpd = dL_ddepth_px[pid], dpa = dL_dalpha_px[pid];
				/* entries at 0-based list index >= n_contrib are skipped (backward.cu:527-529) */
				for (int64_t k = lo + last_contributor - 1; k >= lo; k--) {
					const uint32_t g = st->point_list[k];
					const float dx = st->xy[2 * g] - pxf, dy = st->xy[2 * g + 1] - pyf;
					const float *co = st->conic_op + 4 * g;
					const float power = -0.5f * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
					if (power > 0.0f) continue;
					const float G = expf(power);
					const float alpha = fminf(0.99f, co[3] * G);
					if (alpha < 1.0f / 255.0f) continue;
					T = T / (1.f - alpha);
					const float w = alpha * T;
					double *a = acc + (size_t)g * NA;
					float dL_dopa = 0.0f;
					for (int ch = 0; ch < 3; ch++) {
						const float c = colors[3 * g + ch];
						accum_rec[ch] = last_alpha * last_color[ch] + (1.f - last_alpha) * accum_rec[ch];
						last_color[ch] = c;
						dL_dopa += (c - accum_rec[ch]) * dpix[ch];
						datomic(a + 7 + ch, (double)(w * dpix[ch]));
					}
					for (int ch = 0; ch < S; ch++) {
						const float s = semantics[(size_t)g * S + ch];
						accum_sem[ch] = last_alpha * last_sem[ch] + (1.f - last_alpha) * accum_sem[ch];
						last_sem[ch] = s;
						const float dch = dL_dsem_px[ch * HW + pid];
						dL_dopa += (s - accum_sem[ch]) * dch;
						datomic(a + 11 + ch, (double)(w * dch));
					}
					const float cd = st->depth[g];
					accum_depth = last_alpha * last_depth + (1.f - last_alpha) * accum_depth;
					last_depth = cd;
					dL_dopa += (cd - accum_depth) * dpd;
					datomic(a + 10, (double)(w * dpd));
					accum_alpha = last_alpha + (1.f - last_alpha) * accum_alpha;
					dL_dopa += (1 - accum_alpha) * dpa;
					dL_dopa *= T;
					last_alpha = alpha;
					float bg_dot = 0;
					for (int ch = 0; ch < 3; ch++) bg_dot += pm->bg[ch] * dpix[ch];
					dL_dopa += (-T_final / (1.f - alpha)) * bg_dot;
					const float dL_dG = co[3] * dL_dopa;
					const float gdx = G * dx, gdy = G * dy;
					const float dG_ddelx = -gdx * co[0] - gdy * co[1];
					const float dG_ddely = -gdy * co[2] - gdx * co[1];
					const float mx_ = dL_dG * dG_ddelx * ddelx_dx, my_ = dL_dG * dG_ddely * ddely_dy;
					datomic(a + 0, (double)mx_);
					datomic(a + 1, (double)my_);
					datomic(a + 2, (double)(fabsf(mx_) + fabsf(my_)));
					datomic(a + 3, (double)(-0.5f * gdx * dx * dL_dG));
					datomic(a + 4, (double)(-0.5f * gdx * dy * dL_dG));
					datomic(a + 5, (double)(-0.5f * gdy * dy * dL_dG));
					datomic(a + 6, (double)(G * dL_dopa));
				}
			}
		free(accum_sem);
	}

	const float fy = H / (2.0f * pm->tanfovy), fx = W / (2.0f * pm->tanfovx);
	const float *view = pm->view, *proj = pm->proj;
#pragma omp parallel for schedule(static)
	for (int i = 0; i < P; i++) {
		const double *a = acc + (size_t)i * NA;
		float m2[3] = {(float)a[0], (float)a[1], (float)a[2]};
		g_means2D[3 * i] = m2[0]; g_means2D[3 * i + 1] = m2[1]; g_means2D[3 * i + 2] = m2[2];
		g_opacity[i] = (float)a[6];
		float dcol[3] = {(float)a[7], (float)a[8], (float)a[9]};
		if (g_colors) { g_colors[3 * i] = dcol[0]; g_colors[3 * i + 1] = dcol[1]; g_colors[3 * i + 2] = dcol[2]; }
		for (int ch = 0; ch < S; ch++) g_sem[(size_t)i * S + ch] = (float)a[11 + ch];
		float *gm = g_means3D + 3 * i;
		gm[0] = gm[1] = gm[2] = 0.f;
		if (g_scales) g_scales[3 * i] = g_scales[3 * i + 1] = g_scales[3 * i + 2] = 0.f;
		if (g_rot) g_rot[4 * i] = g_rot[4 * i + 1] = g_rot[4 * i + 2] = g_rot[4 * i + 3] = 0.f;
		if (g_sh) memset(g_sh + (size_t)i * M * 3, 0, (size_t)M * 3 * 4);
		float dcov[6] = {0, 0, 0, 0, 0, 0};
		if (g_cov3D) memset(g_cov3D + 6 * i, 0, 24);
		if (!(st->radii[i] > 0)) continue;

		/* --- backward.cu:144-274: conic -> cov2D -> cov3D and mean (through J) --- */
		const float *mean = means3D + 3 * i;
		const float *c6 = cov3D_precomp ? cov3D_precomp + 6 * i : st->cov3d + 6 * i;
		const float gc[3] = {(float)a[3], (float)a[4], (float)a[5]};
		Proj2 pj; float abc[3];
		proj_jacobian(mean, pm, fx, fy, &pj);
		cov2d_from(&pj, c6, abc);
		const float ca = abc[0], cb = abc[1], cc = abc[2];
		const float denom = ca * cc - cb * cb;
		float dL_da = 0, dL_db = 0, dL_dc = 0;
		const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
		const float(*A)[3] = pj.A; /* A[i][r] == reference T[r][i] */
		if (denom2inv != 0) {
			dL_da = denom2inv * (-cc * cc * gc[0] + 2 * cb * cc * gc[1] + (denom - ca * cc) * gc[2]);
			dL_dc = denom2inv * (-ca * ca * gc[2] + 2 * ca * cb * gc[1] + (denom - ca * cc) * gc[0]);
			dL_db = denom2inv * 2 * (cb * cc * gc[0] - (denom + 2 * cb * cb) * gc[1] + ca * cb * gc[2]);
			dcov[0] = (A[0][0] * A[0][0] * dL_da + A[0][0] * A[1][0] * dL_db + A[1][0] * A[1][0] * dL_dc);
			dcov[3] = (A[0][1] * A[0][1] * dL_da + A[0][1] * A[1][1] * dL_db + A[1][1] * A[1][1] * dL_dc);
			dcov[5] = (A[0][2] * A[0][2] * dL_da + A[0][2] * A[1][2] * dL_db + A[1][2] * A[1][2] * dL_dc);
			dcov[1] = 2 * A[0][0] * A[0][1] * dL_da + (A[0][0] * A[1][1] + A[0][1] * A[1][0]) * dL_db + 2 * A[1][0] * A[1][1] * dL_dc;
			dcov[2] = 2 * A[0][0] * A[0][2] * dL_da + (A[0][0] * A[1][2] + A[0][2] * A[1][0]) * dL_db + 2 * A[1][0] * A[1][2] * dL_dc;
			dcov[4] = 2 * A[0][2] * A[0][1] * dL_da + (A[0][1] * A[1][2] + A[0][2] * A[1][1]) * dL_db + 2 * A[1][1] * A[1][2] * dL_dc;
		}
		const float V[3][3] = {{c6[0], c6[1], c6[2]}, {c6[1], c6[3], c6[4]}, {c6[2], c6[4], c6[5]}};
		float dA[2][3];
		for (int j = 0; j < 3; j++) {
			const float v0 = A[0][0] * V[j][0] + A[0][1] * V[j][1] + A[0][2] * V[j][2];
			const float v1 = A[1][0] * V[j][0] + A[1][1] * V[j][1] + A[1][2] * V[j][2];
			dA[0][j] = 2 * v0 * dL_da + v1 * dL_db;
			dA[1][j] = 2 * v1 * dL_dc + v0 * dL_db;
		}
		/* dJ_il = sum_j Wglm[l][j] * dA_ij with Wglm[l][j] = view[4j + l] (backward.cu:164-167, 243-246) */
		const float dJ00 = view[0] * dA[0][0] + view[4] * dA[0][1] + view[8] * dA[0][2];
		const float dJ02 = view[2] * dA[0][0] + view[6] * dA[0][1] + view[10] * dA[0][2];
		const float dJ11 = view[1] * dA[1][0] + view[5] * dA[1][1] + view[9] * dA[1][2];
		const float dJ12 = view[2] * dA[1][0] + view[6] * dA[1][1] + view[10] * dA[1][2];
		const float tz = 1.f / pj.t[2], tz2 = tz * tz, tz3 = tz2 * tz;
		const float dtx = pj.gx * -fx * tz2 * dJ02;
		const float dty = pj.gy * -fy * tz2 * dJ12;
		const float dtz = -fx * tz2 * dJ00 - fy * tz2 * dJ11 + (2 * fx * pj.t[0]) * tz3 * dJ02 + (2 * fy * pj.t[1]) * tz3 * dJ12;
		/* transformVec4x3Transpose (auxiliary.h:89-97); the reference ASSIGNS here (backward.cu:273) */
		gm[0] = view[0] * dtx + view[1] * dty + view[2] * dtz;
		gm[1] = view[4] * dtx + view[5] * dty + view[6] * dtz;
		gm[2] = view[8] * dtx + view[9] * dty + view[10] * dtz;

		/* --- backward.cu:346-412 --- */
		float mh[4];
		xform4x4(mean, proj, mh);
		const float m_w = 1.0f / (mh[3] + 0.0000001f);
		const float mul1 = (proj[0] * mean[0] + proj[4] * mean[1] + proj[8] * mean[2] + proj[12]) * m_w * m_w;
		const float mul2 = (proj[1] * mean[0] + proj[5] * mean[1] + proj[9] * mean[2] + proj[13]) * m_w * m_w;
		gm[0] += (proj[0] * m_w - proj[3] * mul1) * m2[0] + (proj[1] * m_w - proj[3] * mul2) * m2[1];
		gm[1] += (proj[4] * m_w - proj[7] * mul1) * m2[0] + (proj[5] * m_w - proj[7] * mul2) * m2[1];
		gm[2] += (proj[8] * m_w - proj[11] * mul1) * m2[0] + (proj[9] * m_w - proj[11] * mul2) * m2[1];
		const float ddep = (float)a[10];
		const float mul3 = view[2] * mean[0] + view[6] * mean[1] + view[10] * mean[2] + view[14];
		gm[0] += (view[2] - view[3] * mul3) * ddep;
		gm[1] += (view[6] - view[7] * mul3) * ddep;
		gm[2] += (view[10] - view[11] * mul3) * ddep;

		/* --- SH backward: backward.cu:20-139 --- */
		if (shs) {
			const float *sh = shs + (size_t)i * M * 3;
			float *dsh = g_sh + (size_t)i * M * 3;
			const float dor[3] = {mean[0] - pm->campos[0], mean[1] - pm->campos[1], mean[2] - pm->campos[2]};
			const float len = sqrtf(dor[0] * dor[0] + dor[1] * dor[1] + dor[2] * dor[2]);
			const float x = dor[0] / len, y = dor[1] / len, z = dor[2] / len;
			float dRGB[3];
			for (int c = 0; c < 3; c++) dRGB[c] = dcol[c] * (st->clamped[3 * i + c] ? 0.f : 1.f);
			float ddir[3] = {0, 0, 0};
			const int deg = pm->D;
			for (int c = 0; c < 3; c++) {
#define SHC(k) sh[(k) * 3 + c]
#define DSH(k) dsh[(k) * 3 + c]
				float dx_ = 0, dy_ = 0, dz_ = 0;
				DSH(0) = C0 * dRGB[c];
				if (deg > 0) {
					DSH(1) = (-C1 * y) * dRGB[c]; DSH(2) = (C1 * z) * dRGB[c]; DSH(3) = (-C1 * x) * dRGB[c];
					dx_ = -C1 * SHC(3); dy_ = -C1 * SHC(1); dz_ = C1 * SHC(2);
					if (deg > 1) {
						const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
						DSH(4) = (C2[0] * xy) * dRGB[c]; DSH(5) = (C2[1] * yz) * dRGB[c];
						DSH(6) = (C2[2] * (2.f * zz - xx - yy)) * dRGB[c];
						DSH(7) = (C2[3] * xz) * dRGB[c]; DSH(8) = (C2[4] * (xx - yy)) * dRGB[c];
						dx_ += C2[0] * y * SHC(4) + C2[2] * 2.f * -x * SHC(6) + C2[3] * z * SHC(7) + C2[4] * 2.f * x * SHC(8);
						dy_ += C2[0] * x * SHC(4) + C2[1] * z * SHC(5) + C2[2] * 2.f * -y * SHC(6) + C2[4] * 2.f * -y * SHC(8);
						dz_ += C2[1] * y * SHC(5) + C2[2] * 2.f * 2.f * z * SHC(6) + C2[3] * x * SHC(7);
						if (deg > 2) {
							DSH(9) = (C3[0] * y * (3.f * xx - yy)) * dRGB[c];
							DSH(10) = (C3[1] * xy * z) * dRGB[c];
							DSH(11) = (C3[2] * y * (4.f * zz - xx - yy)) * dRGB[c];
							DSH(12) = (C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy)) * dRGB[c];
							DSH(13) = (C3[4] * x * (4.f * zz - xx - yy)) * dRGB[c];
							DSH(14) = (C3[5] * z * (xx - yy)) * dRGB[c];
							DSH(15) = (C3[6] * x * (xx - 3.f * yy)) * dRGB[c];
							dx_ += (C3[0] * SHC(9) * 3.f * 2.f * xy + C3[1] * SHC(10) * yz + C3[2] * SHC(11) * -2.f * xy +
							        C3[3] * SHC(12) * -3.f * 2.f * xz + C3[4] * SHC(13) * (-3.f * xx + 4.f * zz - yy) +
							        C3[5] * SHC(14) * 2.f * xz + C3[6] * SHC(15) * 3.f * (xx - yy));
							dy_ += (C3[0] * SHC(9) * 3.f * (xx - yy) + C3[1] * SHC(10) * xz +
							        C3[2] * SHC(11) * (-3.f * yy + 4.f * zz - xx) + C3[3] * SHC(12) * -3.f * 2.f * yz +
							        C3[4] * SHC(13) * -2.f * xy + C3[5] * SHC(14) * -2.f * yz + C3[6] * SHC(15) * -3.f * 2.f * xy);
							dz_ += (C3[1] * SHC(10) * xy + C3[2] * SHC(11) * 4.f * 2.f * yz +
							        C3[3] * SHC(12) * 3.f * (2.f * zz - xx - yy) + C3[4] * SHC(13) * 4.f * 2.f * xz +
							        C3[5] * SHC(14) * (xx - yy));
						}
					}
				}
#undef SHC
#undef DSH
				ddir[0] += dx_ * dRGB[c]; ddir[1] += dy_ * dRGB[c]; ddir[2] += dz_ * dRGB[c];
			}
			/* dnormvdv: auxiliary.h:107-117 */
			const float sum2 = dor[0] * dor[0] + dor[1] * dor[1] + dor[2] * dor[2];
			const float inv32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
			gm[0] += ((+sum2 - dor[0] * dor[0]) * ddir[0] - dor[1] * dor[0] * ddir[1] - dor[2] * dor[0] * ddir[2]) * inv32;
			gm[1] += (-dor[0] * dor[1] * ddir[0] + (sum2 - dor[1] * dor[1]) * ddir[1] - dor[2] * dor[1] * ddir[2]) * inv32;
			gm[2] += (-dor[0] * dor[2] * ddir[0] - dor[1] * dor[2] * ddir[1] + (sum2 - dor[2] * dor[2]) * ddir[2]) * inv32;
		}

		/* --- cov3D -> scale / raw quaternion: backward.cu:278-341 (no normalisation Jacobian) --- */
		if (g_cov3D) memcpy(g_cov3D + 6 * i, dcov, 24);
		if (scales) {
			const float *q = rotations + 4 * i;
			const float r = q[0], x = q[1], y = q[2], z = q[3];
			float Rm[3][3];
			rot_from_quat(q, Rm);
			const float s[3] = {pm->scale_modifier * scales[3 * i], pm->scale_modifier * scales[3 * i + 1],
			                    pm->scale_modifier * scales[3 * i + 2]};
			/* M_{k,i} = s_k Rstd_{i,k};  dL/dM = 2 M G with G symmetric (off-diagonals halved) */
			const float Gm[3][3] = {{dcov[0], 0.5f * dcov[1], 0.5f * dcov[2]},
			                        {0.5f * dcov[1], dcov[3], 0.5f * dcov[4]},
			                        {0.5f * dcov[2], 0.5f * dcov[4], dcov[5]}};
			float dM[3][3]; /* dM[k][j] = dL/dM_{k,j} */
			for (int k = 0; k < 3; k++)
				for (int j = 0; j < 3; j++) {
					const float m0 = 2.0f * (s[k] * Rm[0][k]), m1 = 2.0f * (s[k] * Rm[1][k]), m2_ = 2.0f * (s[k] * Rm[2][k]);
					dM[k][j] = m0 * Gm[0][j] + m1 * Gm[1][j] + m2_ * Gm[2][j];
				}
			/* dL/ds_k = <column k of Rstd, row k of dL/dM>  (glm: dot(Rt[k], dL_dMt[k])) */
			for (int k = 0; k < 3; k++)
				g_scales[3 * i + k] = Rm[0][k] * dM[k][0] + Rm[1][k] * dM[k][1] + Rm[2][k] * dM[k][2];
			/* glm dL_dMt[k][j] (column k, row j of dL_dM^T) == dM[k][j]; then scaled by s_k */
			float D_[3][3];
			for (int k = 0; k < 3; k++)
				for (int j = 0; j < 3; j++) D_[k][j] = dM[k][j] * s[k];
			g_rot[4 * i + 0] = 2 * z * (D_[0][1] - D_[1][0]) + 2 * y * (D_[2][0] - D_[0][2]) + 2 * x * (D_[1][2] - D_[2][1]);
			g_rot[4 * i + 1] = 2 * y * (D_[1][0] + D_[0][1]) + 2 * z * (D_[2][0] + D_[0][2]) + 2 * r * (D_[1][2] - D_[2][1]) - 4 * x * (D_[2][2] + D_[1][1]);
			g_rot[4 * i + 2] = 2 * x * (D_[1][0] + D_[0][1]) + 2 * r * (D_[2][0] - D_[0][2]) + 2 * z * (D_[1][2] + D_[2][1]) - 4 * y * (D_[2][2] + D_[0][0]);
			g_rot[4 * i + 3] = 2 * r * (D_[0][1] - D_[1][0]) + 2 * x * (D_[2][0] + D_[0][2]) + 2 * y * (D_[1][2] + D_[2][1]) - 4 * z * (D_[1][1] + D_[0][0]);
		}
	}
	free(acc);
	return 0;
}

/* DGR/rasterizer_impl.cu:54-66 (markVisible) */
void or_mark_visible(int P, const float *means3D, const float *view, uint8_t *present) {
	for (int i = 0; i < P; i++) {
		float pv[3];
		xform4x3(means3D + 3 * i, view, pv);
		present[i] = pv[2] > 0.2f;
	}
}

/* simple-knn distCUDA2 restated: submodules/simple-knn/simple_knn.cu:147-183 computes, for every point, the
 * mean of the squared distances to its 3 nearest OTHER points (exact: the Morton/box machinery is only an
 * acceleration structure; `reject` is an upper bound of the true 3rd-NN distance).  Brute force here. */
void or_knn_mean_dist2(int P, const float *pts, float *out) {
#pragma omp parallel for schedule(static)
	for (int i = 0; i < P; i++) {
		float best[3] = {3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f};
		for (int j = 0; j < P; j++) {
			if (j == i) continue;
			const float dx = pts[3 * j] - pts[3 * i], dy = pts[3 * j + 1] - pts[3 * i + 1], dz = pts[3 * j + 2] - pts[3 * i + 2];
			float d = dx * dx + dy * dy + dz * dz;
			for (int k = 0; k < 3; k++)
				if (best[k] > d) { float t = best[k]; best[k] = d; d = t; }
		}
		out[i] = (best[0] + best[1] + best[2]) / 3.0f;
	}
}

int or_num_threads(void) {
#ifdef _OPENMP
	return omp_get_max_threads();
#else
	return 1;
#endif
}
void or_set_num_threads(int n) {
#ifdef _OPENMP
	omp_set_num_threads(n);
#else
	(void)n;
#endif
}
