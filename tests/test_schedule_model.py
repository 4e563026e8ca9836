"""Happens-before model of ms_factor's stream/event orchestration (csrc/ms_api.cu) and of the
peer-memory panel transport (csrc/ms_p2p.cu).

The factorisation enqueues its kernels on three streams per rank (panel chain, "U1" row stream, main
stream), ordered by events (ev_panel, ev_strip, ev_rest, ev_u2a, ev_free), rotates three panel buffers
and, across ranks, orders the pulls with ready flags and credits.  No GPU is needed to check that this
ordering is SUFFICIENT: the model below replays the enqueue loop of ms_factor for every rank, records
for each kernel the resources it reads and writes (64-row pieces of block rows, panel buffers), builds
the happens-before graph (stream order + event edges + flag/credit edges) and demands that
  * every pair of conflicting accesses (write/write, read/write) is ordered,
  * every piece receives the updates of panels 0 .. J-1 in ascending order before its diag kernel runs
    (what makes the result independent of timing: "scheduling must not change arithmetic"),
  * a peer never pulls a panel before its owner finished it, and no panel buffer is overwritten while a
    local update or a remote pull still reads it.
The loop mirrors ms_factor line by line; when the orchestration changes, this file changes with it.
It has already paid for itself once: it showed that the pre-update pivot test of a block row's later
pieces was reading entries that the look-ahead update of the previous panel could still be writing.
"""
import itertools

import pytest

NPB = 3


class Graph:
    def __init__(self):
        self.nodes = []          # (rank, stream, label, reads, writes)
        self.succ = []
        self.last = {}           # (rank, stream) -> node id
        self.named = {}
        self.deferred = []       # (name of source node, target node id)

    def node(self, rank, stream, label, reads=(), writes=()):
        nid = len(self.nodes)
        self.nodes.append((rank, stream, label, frozenset(reads), frozenset(writes)))
        self.succ.append(set())
        prev = self.last.get((rank, stream))
        if prev is not None:
            self.succ[prev].add(nid)
        self.last[(rank, stream)] = nid
        return nid

    def record(self, rank, stream):
        """cudaEventRecord: the event completes when everything enqueued so far on the stream has."""
        if (rank, stream) not in self.last:
            self.node(rank, stream, "start")
        return self.last[(rank, stream)]

    def wait(self, rank, stream, event_node, label="wait"):
        nid = self.node(rank, stream, label)
        self.succ[event_node].add(nid)
        return nid

    def wait_named(self, rank, stream, name, label):
        nid = self.node(rank, stream, label)
        self.deferred.append((name, nid))
        return nid

    def resolve(self):
        for name, nid in self.deferred:
            self.succ[self.named[name]].add(nid)

    def reach(self):
        n = len(self.nodes)
        order, seen = [], [False] * n
        for s in range(n):                      # iterative DFS post-order
            if seen[s]:
                continue
            stack = [(s, iter(self.succ[s]))]
            seen[s] = True
            while stack:
                v, it = stack[-1]
                for w in it:
                    if not seen[w]:
                        seen[w] = True
                        stack.append((w, iter(self.succ[w])))
                        break
                else:
                    order.append(v)
                    stack.pop()
        r = [0] * n
        for v in order:                         # successors are finished before v
            bits = 1 << v
            for w in self.succ[v]:
                bits |= r[w]
            r[v] = bits
        return r


def build(nb, pieces, P, lookahead=True):
    """Replay of ms_factor (ms_api.cu) for all ranks.  Resources: ("row", J, p) = piece p of block row J,
    ("pb", rank, b, k) = slice k of panel buffer b of a rank (the transport moves a panel as P-1 slices:
    scatter from the owner, all-gather among the receivers — ms_p2p.cu)."""
    g = Graph()
    p2p = P > 1
    nsl = max(P - 1, 1)
    info = {"diag": {}, "update": {}, "pull": {}, "pull2": {}, "panel_done": {}}
    whole = lambda rank, pb: [("pb", rank, pb, k) for k in range(nsl)]
    for rank in range(P):
        sp, su, sm = ("sp", "su", "sm") if lookahead else ("sm", "sm", "sm")
        mine = lambda J: J < nb and J % P == rank
        ev_panel, ev_strip, ev_rest, ev_u2a, ev_free = {}, {}, {}, {}, {}

        def update(stream, I, pb, J, plist):
            nid = g.node(rank, stream, f"upd p{I}->r{J}{plist}", reads=whole(rank, pb),
                         writes=[("row", J, p) for p in plist])
            for p in plist:
                info["update"].setdefault((J, p), []).append((I, nid))
            return nid

        if lookahead:
            ev_strip[0] = g.record(rank, sm)
            g.wait(rank, sp, ev_strip[0])
            g.wait(rank, su, ev_strip[0])
        for I in range(nb):
            pb, owner = I % NPB, I % P
            if lookahead and I >= NPB:
                g.wait(rank, sp, ev_free[I - NPB], f"wait free{I - NPB}")
            if p2p and I >= NPB:                                     # ms_p2p_before_overwrite
                J, jo = I - NPB, (I - NPB) % P
                sliced = J < nb - 1 and P > 2
                if jo == rank or sliced:
                    for r in range(P):
                        if r != rank and (jo == rank or r != jo):
                            g.wait_named(rank, sp, ("credit", J, r, rank), f"wait credit p{J} from {r}")
            if owner == rank:
                if lookahead and I > 0:
                    g.wait(rank, sp, ev_strip[I], f"wait strip{I}")
                for t in range(pieces):                              # factor_block_row
                    # the first piece's diag kernel also runs the reference's pre-update pivot test
                    # (src/solver.cpp:117).  Without look-ahead (or for block row 0) it covers the diagonal
                    # of the WHOLE block row and therefore reads every piece; with look-ahead the later
                    # pieces are tested by pretest_kernel after the ev_rest wait (this model found the race
                    # of the earlier arrangement, where diag(I, 0) read pieces U1 was still updating)
                    split_pre = lookahead and I > 0 and pieces > 1
                    pre = [("row", I, u) for u in range(1, pieces)] if t == 0 and not split_pre else []
                    d = g.node(rank, sp, f"diag r{I}.{t}", reads=pre, writes=[("row", I, t)])
                    info["diag"][(I, t)] = d
                    if I == nb - 1 and t == pieces - 1:
                        continue                                     # nothing to the right of the last piece
                    g.node(rank, sp, f"trsm r{I}.{t}", writes=[("row", I, t)] + whole(rank, pb))
                    if lookahead and I > 0 and t == 0 and pieces > 1:
                        g.wait(rank, sp, ev_rest[I], f"wait rest{I}")
                        g.node(rank, sp, f"pretest r{I}.1..", reads=[("row", I, u) for u in range(1, pieces)])
                    if t + 1 < pieces:
                        g.node(rank, sp, f"strip r{I}.{t}", reads=whole(rank, pb),
                               writes=[("row", I, u) for u in range(t + 1, pieces)])
            if p2p and owner != rank:                                # ms_p2p_wait_panel
                g.wait_named(rank, sp, ("ready", I, rank), f"wait ready p{I}")
                me = (rank - owner - 1) % P                          # receiver index: my slice
                trailing = I < nb - 1
                pull = g.node(rank, sp, f"pull p{I} slice {me} from owner",
                              reads=[("pb", owner, pb, me)] if trailing else [],
                              writes=[("pb", rank, pb, me)] if trailing else [])
                info["pull"][(I, rank)] = pull
                g.named[("credit", I, rank, owner)] = g.node(rank, sp, f"post credit p{I} to owner")
                if trailing and P > 2:
                    g.named[("slice", I, rank)] = g.node(rank, sp, f"post slice p{I}.{me}")
                    for j in range(1, P - 1):
                        idx = (me + j) % (P - 1)
                        r = (owner + 1 + idx) % P
                        g.wait_named(rank, sp, ("slice", I, r), f"wait slice p{I}.{idx} at {r}")
                        p2 = g.node(rank, sp, f"pull p{I} slice {idx} from {r}", reads=[("pb", r, pb, idx)],
                                    writes=[("pb", rank, pb, idx)])
                        info["pull2"][(I, rank, r)] = p2
                        g.named[("credit", I, rank, r)] = g.node(rank, sp, f"post credit p{I} to {r}")
            ev_panel[I] = g.record(rank, sp)
            info["panel_done"][(I, rank)] = ev_panel[I]
            if lookahead:
                g.wait(rank, sm, ev_panel[I], f"main waits panel{I}")
            if p2p and owner == rank:                                # ms_p2p_send_panel
                for r in range(P):
                    if r != rank:
                        g.wait(rank, f"peer{r}", ev_panel[I])
                        g.named[("ready", I, r)] = g.node(rank, f"peer{r}", f"post ready p{I} to {r}")
            if I + 1 >= nb:
                break
            if mine(I + 1):                                          # U1
                if lookahead:
                    g.wait(rank, su, ev_panel[I])
                    if I >= 1:
                        g.wait(rank, su, ev_u2a[I + 1], f"wait u2a{I + 1}")
                update(su, I, pb, I + 1, [0])
                ev_strip[I + 1] = g.record(rank, su)
                if pieces > 1:
                    update(su, I, pb, I + 1, list(range(1, pieces)))
                ev_rest[I + 1] = g.record(rank, su)
            if mine(I + 2):                                          # U2a
                update(sm, I, pb, I + 2, list(range(pieces)))
                ev_u2a[I + 2] = g.record(rank, sm)
                first_rest = I + 3
            else:
                first_rest = I + 3
            rest = [J for J in range(first_rest, nb) if mine(J)]     # U2b
            if rest:
                nid = g.node(rank, sm, f"upd p{I}->rows{rest}", reads=whole(rank, pb),
                             writes=[("row", J, p) for J in rest for p in range(pieces)])
                for J in rest:
                    for p in range(pieces):
                        info["update"].setdefault((J, p), []).append((I, nid))
            ev_free[I] = g.record(rank, sm)
    g.resolve()
    return g, info


CASES = [(nb, pieces, P) for nb in (1, 2, 3, 4, 5, 7, 9, 13) for pieces in (1, 2, 4) for P in (1, 2, 3, 4, 5, 8)
         if P <= max(nb, 1)]


@pytest.mark.parametrize("nb,pieces,P", CASES)
def test_factor_schedule_orders_every_conflict(nb, pieces, P):
    for lookahead in (True, False):
        g, info = build(nb, pieces, P, lookahead)
        reach = g.reach()
        hb = lambda a, b: bool(reach[a] >> b & 1) and a != b
        # 1. conflicting accesses are ordered
        by_res = {}
        for nid, (_, _, _, reads, writes) in enumerate(g.nodes):
            for res in reads:
                by_res.setdefault(res, []).append((nid, False))
            for res in writes:
                by_res.setdefault(res, []).append((nid, True))
        for res, acc in by_res.items():
            for (a, wa), (b, wb) in itertools.combinations(acc, 2):
                if a != b and (wa or wb):
                    assert hb(a, b) or hb(b, a), (res, g.nodes[a][:3], g.nodes[b][:3], lookahead)
        # 2. every piece gets panels 0 .. J-1 in ascending order, all of them before its diag kernel
        for (J, p), d in info["diag"].items():
            ups = sorted(info["update"].get((J, p), []))
            assert [i for i, _ in ups] == list(range(J)), (J, p, ups)
            for (_, a), (_, b) in zip(ups, ups[1:]):
                assert a == b or hb(a, b), (J, p)
            if ups:
                assert hb(ups[-1][1], d), (J, p)
        # 3. a pull starts after the owner's whole panel chain of that block row; a second-phase pull after the
        #    serving receiver's own pull of that slice
        for (I, rank), pull in info["pull"].items():
            assert hb(info["panel_done"][(I, I % P)], pull), (I, rank)
        for (I, rank, src), pull in info["pull2"].items():
            assert hb(info["pull"][(I, src)], pull), (I, rank, src)


def test_model_detects_a_missing_slice_flag():
    """The checker is not vacuous: without the "slice is here" flag a receiver's second-phase pull races with
    the serving receiver's own pull of that slice out of the owner.  (The credits, by contrast, have become
    redundant for P > 2: finishing panel I+1 anywhere needs every receiver's slice of it, hence every receiver's
    completed step I — removing them leaves every conflict ordered, which the model confirms below.)"""
    P = 5
    g, info = build(11, 2, P)
    for name, nid in g.deferred:
        if name[0] == "slice":
            g.succ[g.named[name]].discard(nid)
    reach = g.reach()
    hb = lambda a, b: bool(reach[a] >> b & 1) and a != b
    races = sum(1 for (I, rank, src), p2 in info["pull2"].items()
                if not hb(info["pull"][(I, src)], p2))
    assert races > 0
    g, info = build(11, 2, P)
    for name, nid in g.deferred:
        if name[0] == "credit":
            g.succ[g.named[name]].discard(nid)
    reach = g.reach()
    hb = lambda a, b: bool(reach[a] >> b & 1) and a != b
    for (I, rank, src), p2 in info["pull2"].items():
        pb, idx = I % NPB, ((src - I % P - 1) % P)
        for nid, (r, _, label, _, writes) in enumerate(g.nodes):
            if r == src and ("pb", src, pb, idx) in writes and nid != info["pull"][(I, src)]:
                assert hb(nid, p2) or hb(p2, nid), (I, rank, src, label)
