"""The buffer protocol of the pull-mode multi-GPU reduction (parallel.sharded_backward with GSR_PEER_REDUCE=3,
gsr_backward_partials_marked / gsr_backward_finalize_pull), modelled on the CPU with numpy: N ranks, two alternating buffer
sets, per pass  [composite: local sums + marks] -> barrier -> [chain rule: gather the marked rows of all ranks in rank order,
clear the PREVIOUS pass's marked rows and marks].  Checks, over many passes with random sparse touches (including passes that
touch nothing and Gaussians touched by several ranks):
  * every rank obtains exactly the sum of all ranks' contributions of THIS pass (no leftovers of earlier passes),
  * the result is bit-identical on all ranks (same summation order),
  * the set a composite writes into is all-zero when it starts (accum_is_zero contract),
  * a buffer is never cleared while another rank may still read it (reads of set k happen in the pass that wrote it,
    the clear happens in the following pass, behind that pass's barrier)."""
import numpy as np
import pytest


class Rank:
    def __init__(self, P):
        self.accum = [np.zeros((P, 12), np.float32) for _ in range(2)]
        self.marks = [np.zeros(P, np.uint8) for _ in range(2)]
        self.k = 0


def run_pass(ranks, contributions):
    """contributions[r]: dict {gaussian index: 12-vector} of rank r's band.  Returns the per-rank complete sums."""
    N, P = len(ranks), ranks[0].accum[0].shape[0]
    ks = [rk.k for rk in ranks]
    assert len(set(ks)) == 1, "ranks alternate their buffer sets in lock step"
    k = ks[0]
    # composite (gsr_backward_partials_marked): local adds + marks; the set must be clean
    for r, rk in enumerate(ranks):
        assert not rk.accum[k].any() and not rk.marks[k].any(), "the set a composite writes into must be all-zero"
        for g, v in contributions[r].items():
            rk.accum[k][g] += v
            rk.marks[k][g] = 1
        rk.k ^= 1
    # ---- barrier: every rank's sums and marks of this pass are complete; every rank finished the previous pass's chain rule
    out = []
    for r, rk in enumerate(ranks):          # chain rule (gsr_backward_finalize_pull), any interleaving between ranks is allowed:
        total = np.zeros((P, 12), np.float32)   # it only READS set k of every rank and WRITES its own set k ^ 1
        for q in range(N):
            rows = np.ones(P, bool) if q == r else ranks[q].marks[k].astype(bool)
            total[rows] += ranks[q].accum[k][rows]
        out.append(total)
        prev = rk.marks[k ^ 1].astype(bool)
        rk.accum[k ^ 1][prev] = 0
        rk.marks[k ^ 1][prev] = 0
    return out


@pytest.mark.parametrize("N", [2, 3, 8])
def test_pull_protocol_over_many_passes(N):
    rng = np.random.default_rng(N)
    P = 500
    ranks = [Rank(P) for _ in range(N)]
    for it in range(12):
        contributions = []
        for r in range(N):
            n = 0 if (it == 5) else int(rng.integers(0, 60))            # pass 5 touches nothing at all
            idx = rng.choice(P, size=n, replace=False)
            contributions.append({int(g): rng.normal(size=12).astype(np.float32) for g in idx})
        if it == 7:                                                      # one Gaussian straddling every band
            for r in range(N):
                contributions[r][3] = rng.normal(size=12).astype(np.float32)
        out = run_pass(ranks, contributions)
        want = np.zeros((P, 12), np.float32)
        for r in range(N):                                               # rank order, like the kernel
            for g, v in contributions[r].items():
                want[g] += v
        for r in range(N):
            assert np.array_equal(out[r], out[0]), "ranks disagree"
            assert np.array_equal(out[r], want), f"pass {it}: rank {r} did not obtain this pass's sums"
    # after the last pass only the last set still carries data (cleared by the next pass), the other one is clean
    for rk in ranks:
        assert not rk.accum[rk.k].any() and not rk.marks[rk.k].any()
