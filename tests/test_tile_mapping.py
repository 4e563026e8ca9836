"""CPU emulation of the index arithmetic of the DMMA tile kernel (csrc/ms_kernels_tile.cu).

The kernel cannot run here (no GPU), but its correctness rests on three pieces of pure index
arithmetic that can: the 128-byte TMA swizzle, the lane -> (k-row, column pair) fragment mapping,
and the epilogue's lane -> (row, 4 contiguous columns) ownership.  This test re-derives them from
the formulas in the kernel and checks the claims DESIGN.md makes:
  * every quarter-warp's 16-byte shared-memory loads hit 8 distinct 16-byte slots (no bank conflict);
  * the fragments fed to mma.sync.m8n8k4.f64 and the epilogue mapping together compute C = A^T B;
  * the chunk-major panel-buffer index is a bijection and keeps a column range contiguous.
"""
import numpy as np

KC, CHUNK, BM, BN = 16, 16, 64, 128


def swizzled_byte(k, col):
    """Where TMA (SWIZZLE_128B, box 16 cols x KC rows of fp64) puts element (k, col) of a chunk."""
    return k * 128 + (((col // 2) ^ (k & 7)) << 4) + (col % 2) * 8


def fragment_offset(lane, g, e):
    """The kernel's `off`: kr = 8g + 2q + e, off = kr*128 + ((p ^ (kr & 7)) << 4)."""
    p, q = lane >> 2, lane & 3
    kr = 8 * g + 2 * q + e
    return kr, kr * 128 + ((p ^ (kr & 7)) << 4)


def test_fragment_loads_are_conflict_free_and_fetch_the_intended_elements():
    for g in range(KC // 8):
        for e in range(2):
            for quarter in range(4):
                slots = set()
                for lane in range(8 * quarter, 8 * quarter + 8):
                    kr, off = fragment_offset(lane, g, e)
                    slots.add((off // 16) % 8)
                    p = lane >> 2
                    # the 16 bytes at `off` are columns 2p and 2p+1 of k-row kr
                    assert off == swizzled_byte(kr, 2 * p)
                    assert off + 8 == swizzled_byte(kr, 2 * p + 1)
                assert len(slots) == 8, (g, e, quarter, slots)
    # the two interleaved k-sets of a k8 group cover its eight rows exactly once
    for g in range(KC // 8):
        rows = sorted(fragment_offset(lane, g, e)[0] for e in range(2) for lane in range(4))
        assert rows == list(range(8 * g, 8 * g + 8))


def test_address_registers_reproduce_the_fragment_offsets():
    """tile_dmma_kernel keeps two per-lane byte offsets (k-set e = 0 / 1) and adds 1024 bytes per k8
    group as an immediate; that must be the address of the fragment map above."""
    for lane in range(32):
        p, q = lane >> 2, lane & 3
        off_e = [(2 * q + e) * 128 + ((p ^ (2 * q + e)) << 4) for e in range(2)]
        for g in range(KC // 8):
            for e in range(2):
                assert off_e[e] + g * 1024 == fragment_offset(lane, g, e)[1]
    # the sign word of k-row 8g + 2q + e, read as sgn_chunk[8g + e] with sgn_chunk = sgn + kb*16 + 2q
    for q in range(4):
        for G in range(4):
            g, e = G >> 1, G & 1
            assert 2 * q + 8 * (G >> 1) + (G & 1) == 8 * g + 2 * q + e


def test_producer_consumer_schedule():
    """Host-side model of the tile kernel's single-thread TMA producer: chunk j is requested before chunk
    j - 2 ... j is consumed (never more than STAGES chunks in flight), every stage is only re-armed after
    the chunk that used it was consumed by the requesting warp itself, and the parities line up."""
    STAGES, AHEAD = 4, 2
    for n_chunks in (4, 8, 16, 64):
        pj = 0
        full_phase = [0] * STAGES        # completed phases of full[s]
        empty_phase = [0] * STAGES       # completed phases of empty[s] (as seen once all 8 warps arrived)
        in_flight = {}
        for it in range(n_chunks):
            while pj < n_chunks and pj <= it + AHEAD:
                ps = pj % STAGES
                want = ((pj >> 2) & 1) ^ 1
                # try_wait.parity(P) succeeds when the phase with parity P has completed, i.e. the
                # barrier's current phase parity differs from P
                assert (empty_phase[ps] & 1) != want, ("producer would block on its own warp", it, pj)
                assert ps not in in_flight
                in_flight[ps] = pj
                pj += 1
            s = it % STAGES
            assert in_flight.get(s) == it, "consumer waits for a chunk that was never requested"
            assert (full_phase[s] & 1) == ((it >> 2) & 1)   # waits for the phase with parity (it>>2)&1
            full_phase[s] += 1                               # ... which the TMA bytes complete
            del in_flight[s]
            empty_phase[s] += 1                              # all warps leave the stage
        assert pj == n_chunks and not in_flight


def dmma_8x8x4(c, a, b):
    """mma.sync.aligned.m8n8k4.row.col.f64: lane holds a = A[lane>>2][lane&3], b = B[lane&3][lane>>2],
    c0/c1 = C[lane>>2][(lane&3)*2 + {0,1}]."""
    A = np.zeros((8, 4))
    B = np.zeros((4, 8))
    for lane in range(32):
        A[lane >> 2, lane & 3] = a[lane]
        B[lane & 3, lane >> 2] = b[lane]
    D = A @ B
    for lane in range(32):
        for t in range(2):
            c[lane, t] += D[lane >> 2, (lane & 3) * 2 + t]


def test_warp_tile_computes_At_B_through_the_kernel_mappings():
    rng = np.random.default_rng(0)
    K = 32
    PA = rng.standard_normal((K, BM))      # panel columns that match the tile's rows
    PB = rng.standard_normal((K, BN))      # panel columns that match the tile's columns
    want = PA.T @ PB
    got = np.full((BM, BN), np.nan)
    for warp in range(8):
        wm, wn = warp >> 2, warp & 3
        acc = np.zeros((2, 2, 2, 2, 32, 2))          # [ib][jb][a][b][lane][e]
        for kb in range(K // KC):
            for g in range(KC // 8):
                for e in range(2):
                    fa = np.zeros((2, 32, 2))
                    fb = np.zeros((2, 32, 2))
                    for lane in range(32):
                        p = lane >> 2
                        kr, _ = fragment_offset(lane, g, e)
                        k = kb * KC + kr
                        for ib in range(2):
                            col = (wm * 2 + ib) * CHUNK + 2 * p
                            fa[ib, lane] = PA[k, col], PA[k, col + 1]
                        for jb in range(2):
                            col = (wn * 2 + jb) * CHUNK + 2 * p
                            fb[jb, lane] = PB[k, col], PB[k, col + 1]
                    for ib in range(2):
                        for jb in range(2):
                            for a in range(2):
                                for b in range(2):
                                    dmma_8x8x4(acc[ib, jb, a, b], fa[ib, :, a], fb[jb, :, b])
        # epilogue: lane owns row wm*32 + 16*ib + 2p + a, columns wn*32 + 16*jb + 4q + {0,1,2,3}
        # with values acc[a][b=0][e=0], acc[a][1][0], acc[a][0][1], acc[a][1][1]
        for lane in range(32):
            p, q = lane >> 2, lane & 3
            for ib in range(2):
                for a in range(2):
                    r = wm * 32 + 16 * ib + 2 * p + a
                    for jb in range(2):
                        c0 = wn * 32 + 16 * jb + 4 * q
                        vals = [acc[ib, jb, a, 0, lane, 0], acc[ib, jb, a, 1, lane, 0],
                                acc[ib, jb, a, 0, lane, 1], acc[ib, jb, a, 1, lane, 1]]
                        for t in range(4):
                            assert np.isnan(got[r, c0 + t]), "two lanes own the same element"
                            got[r, c0 + t] = vals[t]
    assert not np.isnan(got).any(), "some element of the tile is owned by no lane"
    assert np.allclose(got, want, rtol=1e-13, atol=1e-13)


def chunk_major_index(k, j, H):
    return ((j // CHUNK) * H + k) * CHUNK + (j % CHUNK)


def test_chunk_major_panel_buffer_layout():
    H, ld = 128, 96
    seen = set()
    for k in range(H):
        for j in range(ld):
            seen.add(chunk_major_index(k, j, H))
    assert seen == set(range(H * ld))                       # bijection onto the buffer
    # the columns >= c0 (c0 a multiple of 16) of ALL k rows form one contiguous tail: what one
    # ncclBroadcast / one peer copy of step I moves
    c0 = 32
    tail = sorted(chunk_major_index(k, j, H) for k in range(H) for j in range(c0, ld))
    assert tail == list(range((c0 // CHUNK) * H * CHUNK, H * ld))
    # a TMA box (16 columns x 16 k-rows) is one contiguous 2 KB run
    box = sorted(chunk_major_index(k, j, H) for k in range(16, 32) for j in range(48, 64))
    assert box == list(range(box[0], box[0] + 256))


def test_rank3_tensor_map_addresses_the_chunk_major_buffer():
    """The tile kernel views a panel buffer through a rank-3 tensor map: dims {16, H, ld/16}, strides
    {128 B, H*128 B}; coordinate (x, y, z) = (column in chunk, k row, chunk).  That must be the element
    chunk_major_index(k = y, j = 16 z + x), and a box {16, 16, 4} must land as four consecutive 2 KB
    chunk images (x fastest, then k, then chunk) — the layout the fragment loads assume."""
    H, ld = 192, 160
    for z in range(ld // CHUNK):
        for y in (0, 1, 17, H - 1):
            for x in range(CHUNK):
                byte = x * 8 + y * 128 + z * H * 128
                assert byte == 8 * chunk_major_index(y, CHUNK * z + x, H)
    # box at (0, k0, c0): element (x, dy, dz) is written at smem offset ((dz*16 + dy)*16 + x)*8 before
    # swizzling = chunk image dz (2 KB each), k row dy, column x
    for dz in range(4):
        for dy in range(KC):
            for x in range(CHUNK):
                linear = ((dz * KC + dy) * CHUNK + x) * 8
                assert linear == dz * 2048 + dy * 128 + x * 8
                # the swizzle only permutes 16-byte slots inside a 128-byte row, so the fragment formula holds
                assert swizzled_byte(dy, x) // 128 == dy


# ---- the tile list of one update launch (tile_row_global / tile_first_in_row / tile_build_bands and
# ---- the walk in tile_dmma_kernel: band table, prefix scan over row tiles, two binary searches, T tiles per CTA)
def _tile_row_global(lrow_begin, m, rank, nranks, rt):
    l = lrow_begin + rt * BM
    lb = l // m
    return (lb * nranks + rank) * m + (l - lb * m)


def _launch_tiles(n, m, rank, nranks, lrow_begin, lrow_end, tiles_per_cta, ncols=None, band_w_min=64, max_bands=64):
    """tile_build_bands (host) + tile_decode (device) of csrc/ms_kernels_tile.cu: the stored tiles are
    walked band by band (band = band_w column tiles), row tile by row tile inside a band."""
    ncols = n if ncols is None else ncols                # n + 1 with the augmented column
    n_rt = (lrow_end - lrow_begin + BM - 1) // BM
    total_ct = (ncols + BN - 1) // BN
    first = []
    for rt in range(n_rt):
        g = _tile_row_global(lrow_begin, m, rank, nranks, rt)
        first.append(min(g // BN, total_ct) if g < n else total_ct)
    assert first == sorted(first)                        # the band table relies on it
    prefF = [0]
    for f in first:
        prefF.append(prefF[-1] + f)
    band_w = max(band_w_min, (total_ct + max_bands - 1) // max_bands)
    n_bands = max(1, (total_ct + band_w - 1) // band_w)
    start, r1s, r2s = [0], [], []
    for b in range(n_bands):
        lo, hi = b * band_w, min((b + 1) * band_w, total_ct)
        r1 = sum(1 for f in first if f <= lo)
        r2 = sum(1 for f in first if f < hi)
        r1s.append(r1)
        r2s.append(r2)
        start.append(start[-1] + r1 * (hi - lo) + (r2 - r1) * hi - (prefF[r2] - prefF[r1]))
    total = start[-1]
    assert total == sum(total_ct - f for f in first)
    tiles = []
    n_cta = (total + tiles_per_cta - 1) // tiles_per_cta
    for cta in range(n_cta):
        for t in range(cta * tiles_per_cta, min(total, (cta + 1) * tiles_per_cta)):
            b, bh = 0, n_bands - 1                       # the kernel's two binary searches
            while b < bh:
                mid = (b + bh + 1) >> 1
                if start[mid] <= t:
                    b = mid
                else:
                    bh = mid - 1
            lo_ct, hi_ct = b * band_w, min((b + 1) * band_w, total_ct)
            wfull, r1, r2 = hi_ct - lo_ct, r1s[b], r2s[b]
            o = t - start[b]
            if o < r1 * wfull:
                rt = o // wfull
                ct = lo_ct + (o - rt * wfull)
            else:
                o -= r1 * wfull
                lo, hi = r1, r2 - 1
                while lo < hi:
                    mid = (lo + hi + 1) >> 1
                    if (mid - r1) * hi_ct - (prefF[mid] - prefF[r1]) <= o:
                        lo = mid
                    else:
                        hi = mid - 1
                rt = lo
                ct = (prefF[lo + 1] - prefF[lo]) + (o - ((lo - r1) * hi_ct - (prefF[lo] - prefF[r1])))
            assert lo_ct <= ct < hi_ct and ct >= first[rt]
            tiles.append((rt, _tile_row_global(lrow_begin, m, rank, nranks, rt), ct))
    return tiles, total


def test_banded_walk_keeps_neighbouring_tiles_on_one_panel_slab():
    """The point of the bands: the tiles in flight at any time (~1200 consecutive entries of the list) read
    at most one band's worth of panel columns as their B operand plus a few row tiles' A slabs."""
    n, m = 65536, 256
    tiles, total = _launch_tiles(n, m, 0, 1, 768, n, 4, ncols=n + 1)
    assert total == sum((n + 1 + BN - 1) // BN - (g // BN) for g in range(768, n, BM))
    window = 1184
    for s0 in range(0, total - window, 50021):
        cts = {ct for _, _, ct in tiles[s0:s0 + window]}
        rts = {rt for rt, _, _ in tiles[s0:s0 + window]}
        assert max(cts) - min(cts) < 2 * 64           # at most two neighbouring bands
        assert len(rts) <= window // 16 + 2


def test_update_launch_covers_the_stored_trapezoid_exactly_once():
    """Every stored element (j >= i) of the rows an update launch owns is written by exactly one tile;
    nothing left of the row's diagonal BLOCK is ever touched; the A operand's first chunk is the panel
    column that matches the tile's first row."""
    for n, m, nranks in [(640, 128, 1), (1000, 192, 1), (1024, 256, 2), (1111, 64, 3), (2048, 256, 4), (700, 64, 2)]:
        nblocks = (n + m - 1) // m
        for rank in range(nranks):
            mine = [J for J in range(nblocks) if J % nranks == rank]
            for I in (0, 1, nblocks // 2):
                later = [J for J in mine if J > I]
                if not later:
                    continue
                lrow_begin = (later[0] // nranks) * m
                rows_local = sum(min(m, n - J * m) for J in mine)
                lrow_end = rows_local
                for T, bw in ((1, 64), (4, 2), (4, 1), (3, 3)):   # small band widths: several bands at these n
                    tiles, total = _launch_tiles(n, m, rank, nranks, lrow_begin, lrow_end, T, band_w_min=bw)
                    assert len(tiles) == total and len(set(tiles)) == total
                    written = {}
                    for rt, g_row0, ct in tiles:
                        assert g_row0 % CHUNK == 0                     # achunk0 = g_row0 / 16 is exact
                        rows_valid = min(BM, n - g_row0)
                        for r in range(rows_valid):
                            i = g_row0 + r
                            for j in range(max(ct * BN, g_row0), min((ct + 1) * BN, n)):
                                assert (i, j) not in written
                                written[(i, j)] = True
                                assert j >= (i // m) * m               # never left of the diagonal block
                    want = {(i, j) for J in later for i in range(J * m, min((J + 1) * m, n)) for j in range(i, n)}
                    assert want <= set(written), (n, m, nranks, rank, I, T)
                    # what is written beyond the stored triangle is only the lower part of diagonal tiles
                    for (i, j) in set(written) - want:
                        assert j < i and j >= (i // BM) * BM
