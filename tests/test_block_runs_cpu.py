"""The bookkeeping of the block-run exchange (street_gaussians_b200/csrc/sgr_common.cuh, include/sgr.h: sgr_sharded_forward) restated in
numpy and checked for the invariants the CUDA kernels rely on — no GPU needed, no kernel is run here (the kernels themselves are
checked bit-for-bit against the single-GPU render by the `-m gpu` emulated-rank tests and by bench.py's parity_n):

  * owner s, block b (256 consecutive Gaussians) delivers the records rank d needs as ONE run into slots [s*chunk + 256 b, +c(s,b,d));
  * rank d enumerates its delivered Gaussians from the exclusive prefix of its run-length table by binary search (count_runs_kernel);
  * that enumeration is in ascending slot order AND ascending global-id order (the tie order of the depth sort on one GPU);
  * the owner re-derives each record's slot on every destination from the stored masks alone (gather_runs in the backward)."""
import numpy as np
import pytest

RUN = 256


def deliver(masks, rank, world, chunk):
    """What preprocess_fwd_kernel<false, SCATTER> of `rank` does: returns {dest: (slots, global ids)} and its rows of the count tables."""
    nblk = (chunk + RUN - 1) // RUN
    out = {d: ([], []) for d in range(world)}
    cnt = np.zeros((world, nblk), dtype=np.int64)  # cnt[d][b] goes to rank d's table row `rank`
    for b in range(nblk):
        idx = np.arange(b * RUN, min((b + 1) * RUN, len(masks)))
        for d in range(world):
            hit = idx[(masks[idx] >> d) & 1 == 1]           # ascending thread order = ascending Gaussian order
            cnt[d, b] = len(hit)
            out[d][0].extend(rank * chunk + b * RUN + np.arange(len(hit)))
            out[d][1].extend(rank * chunk + hit)
    return out, cnt


def enumerate_delivered(table, world, chunk):
    """run_prefix_kernel + the binary search of count_runs_kernel on one receiving rank; table[s][b] = run lengths."""
    nblk = table.shape[1]
    flat = table.reshape(-1)
    pre = np.concatenate([[0], np.cumsum(flat)[:-1]])
    total = int(flat.sum())
    slots = []
    for j in range(total):
        lo, hi = 0, len(pre)
        while hi - lo > 1:
            mid = (lo + hi) // 2
            if pre[mid] <= j:
                lo = mid
            else:
                hi = mid
        s, b = divmod(lo, nblk)
        slots.append(s * chunk + b * RUN + (j - pre[lo]))
    return np.array(slots, dtype=np.int64), total


@pytest.mark.parametrize("world,chunk,seed", [(1, 700, 0), (2, 1000, 1), (3, 513, 2), (8, 300, 3), (4, 256, 4), (5, 1, 5)])
def test_slot_order_is_global_id_order_and_backward_positions_match(world, chunk, seed):
    rng = np.random.default_rng(seed)
    local_n = [chunk] * (world - 1) + [max(0, chunk - rng.integers(0, min(chunk, 40) + 1))]  # the last rank may own fewer
    masks = []
    for r in range(world):
        m = rng.integers(0, 1 << world, size=local_n[r], dtype=np.int64)
        m[rng.random(local_n[r]) < 0.3] = 0                                                    # culled Gaussians go nowhere
        masks.append(m)
    sent = [deliver(masks[r], r, world, chunk) for r in range(world)]
    nblk = (chunk + RUN - 1) // RUN
    for d in range(world):
        table = np.stack([sent[s][1][d] for s in range(world)])                                 # [s][b], as the owners filled it
        slots, total = enumerate_delivered(table, world, chunk)
        # what actually sits in those slots: the owners' stores
        slot_to_gid = {}
        for s in range(world):
            sl, gid = sent[s][0][d]
            assert len(set(sl)) == len(sl)
            slot_to_gid.update(zip(sl, gid))
        assert total == len(slot_to_gid) == sum(int(((masks[s] >> d) & 1).sum()) for s in range(world))
        assert sorted(slot_to_gid) == list(slots)                                               # every delivered slot, in ascending order
        gids = np.array([slot_to_gid[s] for s in slots])
        assert np.all(np.diff(gids) > 0)                                                        # ascending slot order == ascending global id
        assert np.all(slots // chunk == gids // chunk)                                          # a slot stays inside its owner's region
    # backward: the owner re-derives the slot of Gaussian i on destination d from the masks of its block alone
    for r in range(world):
        for d in range(world):
            sl, gid = sent[r][0][d]
            pos = dict(zip(gid, sl))
            for i in np.nonzero((masks[r] >> d) & 1)[0][:: max(1, local_n[r] // 50)]:
                b = i // RUN
                rank_in_block = int(((masks[r][b * RUN: i] >> d) & 1).sum())
                assert pos[r * chunk + i] == r * chunk + b * RUN + rank_in_block
