"""Host-side arithmetic of the block-row cyclic layout (mirrors ms_ctx in csrc/ms_common.cuh).

Used by bench.py for the algorithmic byte / flop accounting and by the CPU tests; no device
code here.
"""


def n_blocks(n, m):
    return n // m + (1 if n % m else 0)


def block_rows(n, m, I):
    return m if (I + 1) * m <= n else n - I * m


def owner(I, nranks):
    return I % nranks


def rows_local(n, m, rank, nranks):
    return sum(block_rows(n, m, I) for I in range(rank, n_blocks(n, m), nranks))


def local_row(n, m, g, nranks):
    I = g // m
    return (I // nranks) * m + (g - I * m)


def ref_alloc_size(n, m, rank=0, nranks=1):
    """Elements the reference allocates on `rank` (src/init.cpp:4-27): block COLUMN J % P."""
    total = 0
    for J in range(rank, n_blocks(n, m), nranks):
        w = block_rows(n, m, J)
        total += J * m * w + w * (w + 1) // 2
    return total


def factor_flops(n):
    """The metric's flop count, n^3/3 (latex/rep.tex:91-92, BASELINE.json)."""
    return n ** 3 / 3.0


def update_flops(n, m, rank=0, nranks=1):
    """Algorithmic flops of all rank-m trailing updates on `rank`: 2*m per stored element of every
    block row J > I, summed over steps I (what stats.update_flops accumulates)."""
    nb = n_blocks(n, m)
    total = 0.0
    for I in range(nb - 1):
        for J in range(I + 1, nb):
            if owner(J, nranks) != rank:
                continue
            rows = block_rows(n, m, J)
            g0 = J * m
            total += 2.0 * m * (rows * (n - g0) - 0.5 * rows * (rows - 1))
    return total


def solve_bytes(n):
    """Each triangular sweep reads every stored element of R once: 8 B * n(n+1)/2."""
    return 4.0 * n * (n + 1)
