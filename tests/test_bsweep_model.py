"""CPU model of the control flow of bsweep_kernel (csrc/ms_kernels_solve.cu): the single-launch backward
sweep that runs on every rank and exchanges x through peer memory.  The model executes the kernel's
per-CTA loop (same index arithmetic: owner_of, lidx_of, lcut, the `ahead` rule, the chunk walk) for every
CTA of every rank and checks what a wrong index would break on the GPU as a hang or a wrong x:
  * every piece is solved exactly once, by the CTA that owns its chunk, with a COMPLETE accumulator
    (every block R(chunk, t') with t' > piece applied exactly once before the solve);
  * a CTA only ever waits for a piece that was published in an earlier iteration or by a CTA that does not
    itself wait in this iteration (no circular wait);
  * no block is applied twice or to the wrong accumulator slot, and slots never exceed SWEEP_MAX_SLOTS.
It is a model of the schedule, not evidence about the kernel's arithmetic — the GPU parity tests are."""
import itertools

import pytest

SWEEP_MAX_SLOTS = 24


def simulate(T, q, P, Gs):
    """T global pieces, q pieces per block row, P ranks, Gs[rank] CTAs on that rank."""
    owner_of = lambda t: (t // q) % P
    lidx_of = lambda t: ((t // q) // P) * q + t % q

    def lcut(t, rank):
        I = t // q
        if I % P == rank:
            return (I // P) * q + t % q
        return (I // P + (1 if I % P > rank else 0)) * q

    def gp(lc, rank):
        return ((lc // q) * P + rank) * q + lc % q

    LT = [sum(1 for t in range(T) if owner_of(t) == r) for r in range(P)]
    applied = {}          # (rank, lc) -> set of pieces applied to that chunk's accumulator
    solved_by = {}        # piece -> (rank, cta, iteration)
    published_at = {}     # piece -> iteration in which its flag is raised
    solved_ahead = {(r, g): False for r in range(P) for g in range(Gs[r])}
    for r in range(P):
        G = Gs[r]
        assert G >= 1 or LT[r] == 0
        if LT[r]:
            assert (LT[r] + G - 1) // G <= SWEEP_MAX_SLOTS
        for lc in range(LT[r]):
            applied[(r, lc)] = set()
            assert owner_of(gp(lc, r)) == r and lidx_of(gp(lc, r)) == lc     # the two maps are inverse

    def solve(t, r, g, it):
        lc = lidx_of(t)
        assert owner_of(t) == r and lc % Gs[r] == g
        assert t not in solved_by, f"piece {t} solved twice"
        assert applied[(r, lc)] == set(range(t + 1, T)), (t, sorted(applied[(r, lc)]))
        solved_by[t] = (r, g, it)
        published_at[t] = it

    for t in range(T - 1, -1, -1):
        tn = t - 1
        plans = []
        for r in range(P):
            G = Gs[r]
            for g in range(G if LT[r] else 0):
                own_next = tn >= 0 and owner_of(tn) == r and lidx_of(tn) % G == g
                mine = owner_of(t) == r and lidx_of(t) % G == g
                ahead = own_next and not (mine and not solved_ahead[(r, g)])
                plans.append((r, g, mine, ahead))
        # phase 1: owners that solve piece t in this very iteration (they wait for nobody)
        for r, g, mine, ahead in plans:
            if mine and not solved_ahead[(r, g)]:
                solve(t, r, g, t)
        assert t in published_at, f"nobody publishes piece {t}"
        assert published_at[t] >= t          # raised in iteration t, or in iteration t + 1 by an `ahead` CTA
        # phase 2: everybody applies x_t; `ahead` CTAs then solve and publish piece t - 1
        for r, g, mine, ahead in plans:
            G = Gs[r]
            solved_ahead[(r, g)] = False
            lc = lcut(t, r) - 1
            if lc >= 0:
                lc -= ((lc - g) % G + G) % G
            if ahead:
                assert lc == lidx_of(tn), (t, r, g, lc, lidx_of(tn))
                applied[(r, lc)].add(t)
                solve(tn, r, g, t)
                solved_ahead[(r, g)] = True
                lc -= G
            while lc >= 0:
                assert gp(lc, r) < t and lc % G == g and lc // G < SWEEP_MAX_SLOTS
                assert t not in applied[(r, lc)], "block applied twice"
                applied[(r, lc)].add(t)
                lc -= G
    assert sorted(solved_by) == list(range(T))
    for (r, lc), s in applied.items():
        assert s == set(range(gp(lc, r) + 1, T))
    return solved_by


@pytest.mark.parametrize("P", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("q", [1, 2, 4])
def test_backward_sweep_schedule(P, q):
    for nblocks in (1, 2, P - 1, P, P + 1, 2 * P + 1, 5 * P + 3, 64):
        if nblocks < 1:
            continue
        for short_last in (0, 1):
            T = nblocks * q - (short_last if q > 1 else 0)     # the last block row may hold fewer pieces (n % m != 0)
            if T < 1:
                continue
            for G in (1, 2, 3, 5, 592):
                LT = [sum(1 for t in range(T) if (t // q) % P == r) for r in range(P)]
                Gs = [max(1, min(G, lt)) for lt in LT]
                if any(lt and (lt + g - 1) // g > SWEEP_MAX_SLOTS for lt, g in zip(LT, Gs)):
                    continue
                simulate(T, q, P, Gs)


def test_backward_sweep_schedule_at_the_headline_size():
    """n = 65536, m = 256: 1024 pieces, 4 per block row, 1 and 8 ranks with the launcher's grid sizes."""
    for P in (1, 8):
        LT = 1024 // P
        sol = simulate(1024, 4, P, [min(LT, 4 * 148)] * P)
        assert len(sol) == 1024
