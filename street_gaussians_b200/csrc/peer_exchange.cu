// peer_exchange.cu — the two exchange steps of the Gaussian-sharded mode done with plain loads/stores on NVLink peer
// memory instead of NCCL collectives (include/sgr.h: sgr_scatter_records, sgr_gather_grad2d).
//
// With cyclic tile-row bands (row r belongs to rank r % world) a Gaussian whose tile rectangle spans h rows is needed by
// min(h, world) ranks only — 1 or 2 for the bulk of a street scene — while an all-gather ships every record to every
// rank.  So the owner WRITES each 48-B record straight into the gathered array of exactly the ranks that need it (and a
// 4-B radius, 0 = "not yours", to everyone), and in backward READS the partial grad2d rows back from exactly those
// ranks.  Arrays stay indexed by the global Gaussian id, so binning order and results are bit-identical to the NCCL
// path; the caller brackets the two steps with device-side barriers (torch symmetric memory).
#include "sgr_common.cuh"

namespace sgr {

__device__ __forceinline__ uint32_t ranks_of(const FrameDev &f, const float4 q0, int radius, int world) {
	int x0, y0, x1, y1;
	tile_rect(q0.x, q0.y, radius, f.gx, f.gy, x0, y0, x1, y1);
	return x1 > x0 ? touched_ranks(y0, y1, world) : 0u;
}

// one thread per slot of this rank's chunk (slots past f.P are padding: radius 0 everywhere)
__global__ void __launch_bounds__(256) scatter_records_kernel(const FrameDev f, const PeerTable pt, const GaussRec *__restrict__ rec,
                                                             const int32_t *__restrict__ radii) {
	const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= pt.chunk) return;
	int r = 0;
	uint32_t mask = 0u;
	GaussRec v;
	v.q0 = v.q1 = v.q2 = make_float4(0.f, 0.f, 0.f, 0.f);
	if (idx < f.P) {
		r = radii[idx];
		if (r > 0) {
			v = rec[idx];
			mask = ranks_of(f, v.q0, r, pt.world);
		}
	}
	const size_t g = (size_t)pt.rank * (size_t)pt.chunk + (size_t)idx;
	for (int p = 0; p < pt.world; p++) {
		const bool hit = (mask >> p) & 1u;
		if (hit) pt.rec[p][g] = v;
		pt.radii[p][g] = hit ? r : 0;
	}
}

// grad2d_local[i] = sum over the ranks that rendered Gaussian i of their partial row (fixed ascending-rank order)
__global__ void __launch_bounds__(256) gather_grad2d_kernel(const FrameDev f, const PeerTable pt, const GaussRec *__restrict__ rec,
                                                           const int32_t *__restrict__ radii, float *__restrict__ out) {
	const int idx = blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= f.P) return;
	float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a, c = a;
	const int r = radii[idx];
	if (r > 0) {
		const uint32_t mask = ranks_of(f, rec[idx].q0, r, pt.world);
		const size_t g = (size_t)pt.rank * (size_t)pt.chunk + (size_t)idx;
		for (int p = 0; p < pt.world; p++) {
			if (!((mask >> p) & 1u)) continue;
			const float4 *src = reinterpret_cast<const float4 *>(pt.grad2d[p] + g * 12);
			const float4 x = src[0], y = src[1], z = src[2];
			a.x += x.x; a.y += x.y; a.z += x.z; a.w += x.w;
			b.x += y.x; b.y += y.y; b.z += y.z; b.w += y.w;
			c.x += z.x; c.y += z.y; c.z += z.z; c.w += z.w;
		}
	}
	float4 *dst = reinterpret_cast<float4 *>(out + (size_t)idx * 12);
	dst[0] = a; dst[1] = b; dst[2] = c;
}

// Cross-rank barrier on the caller's stream: thread p announces this rank's arrival in rank p's pad (a release store at system
// scope, after a system-scope fence that orders this rank's earlier peer stores — issued by previous kernels of the stream —
// before it) and then waits until rank p has announced itself in this rank's pad.  Epochs increase by one per barrier and every
// rank issues the same sequence of barriers, so one pad suffices: a peer that is already one barrier ahead has written a larger
// epoch, which also satisfies the wait.  epoch == 0 ("auto"): the epoch is kept ON THE DEVICE — slot kMaxPeers of the rank's own pad
// counts its barriers — so the launch carries no per-step host value and a captured CUDA graph of a step can be replayed.
// The spin is bounded (2 s) so a missing peer cannot wedge the GPU; a timeout sets status[5] and the frame is garbage.
__global__ void peer_barrier_kernel(const PeerTable pt, uint32_t epoch, uint32_t *__restrict__ status) {
	const int p = threadIdx.x;
	if (epoch == 0u) {
		uint32_t e = 0u;
		if (p == 0) {
			uint32_t *counter = pt.flags[pt.rank] + kMaxPeers;
			e = *counter + 1u;
			*counter = e;
		}
		epoch = __shfl_sync(0xffffffffu, e, 0);
	}
	if (p >= pt.world) return;
	__threadfence_system();
	uint32_t *theirs = pt.flags[p] + pt.rank;
	asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(theirs), "r"(epoch) : "memory");
	const uint32_t *mine = pt.flags[pt.rank] + p;
	unsigned long long t0;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
	for (;;) {
		uint32_t v;
		asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
		if ((int32_t)(v - epoch) >= 0) break;
		__nanosleep(40);
		unsigned long long t1;
		asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
		if (t1 - t0 > 2000000000ull) {
			if (status) atomicExch(&status[5], epoch);
			break;
		}
	}
}
cudaError_t launch_peer_barrier(const PeerTable &pt, uint32_t epoch, uint32_t *status, cudaStream_t st) {
	if (pt.world <= 1) return cudaSuccess;
	count_launch();
	peer_barrier_kernel<<<1, 32, 0, st>>>(pt, epoch, status);
	return cudaGetLastError();
}

cudaError_t launch_scatter_records(const FrameDev &f, const PeerTable &pt, const GaussRec *rec, const int32_t *radii, cudaStream_t st) {
	if (pt.chunk == 0) return cudaSuccess;
	count_launch();
	scatter_records_kernel<<<(unsigned)((pt.chunk + 255) / 256), 256, 0, st>>>(f, pt, rec, radii);
	return cudaGetLastError();
}
cudaError_t launch_gather_grad2d(const FrameDev &f, const PeerTable &pt, const GaussRec *rec, const int32_t *radii, float *out,
                                 cudaStream_t st) {
	if (f.P == 0) return cudaSuccess;
	count_launch();
	gather_grad2d_kernel<<<(f.P + 255) / 256, 256, 0, st>>>(f, pt, rec, radii, out);
	return cudaGetLastError();
}

}  // namespace sgr
