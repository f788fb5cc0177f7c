"""A small executable model of the peer-memory halo protocol of the fused apply kernel (DESIGN.md section 4,
csrc/jp_csr.cu spmv_fused_halo_kernel / csrc/jp_p2p.cuh), explored under random schedules on the CPU.

Per rank and per apply (epoch e), in program order of the kernel's CTA roles; every receiver has TWO halo buffers
and apply e uses buffer e & 1:
  PACK      wait  ack[dst] >= e-2  for every destination   (back-pressure: what buffer e & 1 held has been consumed)
            write halo[dst][me][e & 1] := (me, e);  ready[dst][me] := e
  INTERIOR  (no communication)
  BOUNDARY  wait  ready[me][src] >= e  for every source
            read  halo[me][src][e & 1]  -> must be exactly (src, e)
            ack[src][me] := e
Kernels of one rank run in stream order (apply e+1 starts after apply e has finished), ranks are asynchronous.
The model checks, for random interleavings and random neighbour graphs, that (1) every rank finishes every epoch
(no deadlock), (2) a boundary phase always reads the halo of ITS epoch (never a stale or a too-new one), i.e. no
sender overwrites a buffer that has not been consumed, and (3) that a rank really can run a whole apply ahead of a
neighbour (the point of the second buffer: neighbours are not lock-stepped)."""
import random

import pytest


def run(P, sends, epochs, rng):
    # sends[p] = destinations of rank p; sources[q] = ranks that send to q
    sources = {q: [p for p in range(P) if q in sends[p]] for q in range(P)}
    ready = {(q, p): 0 for p in range(P) for q in sends[p]}   # ready[q][p]: written by p, read by q
    ack = {(p, q): 0 for p in range(P) for q in sends[p]}     # ack[p][q]: written by q, read by p
    halo = {(q, p): [None, None] for p in range(P) for q in sends[p]}
    max_lead = 0
    # per-rank program counter: (epoch, phase) with phase 0 = pack-wait, 1 = pack-write, 2 = boundary-wait, 3 = boundary-read
    pc = {p: [1, 0] for p in range(P)}
    steps = 0
    while any(pc[p][0] <= epochs for p in range(P)):
        runnable = []
        for p in range(P):
            e, ph = pc[p]
            if e > epochs:
                continue
            if ph == 0 and all(ack[(p, q)] >= e - 2 for q in sends[p]):
                runnable.append(p)
            elif ph == 1 or ph == 3:
                runnable.append(p)
            elif ph == 2 and all(ready[(p, s)] >= e for s in sources[p]):
                runnable.append(p)
        assert runnable, f"deadlock at {pc}"
        p = rng.choice(runnable)
        e, ph = pc[p]
        if ph == 0:
            pc[p][1] = 1
        elif ph == 1:
            for q in sends[p]:
                # the buffer must have been consumed: its last content is from epoch e-2 and was acknowledged
                assert halo[(q, p)][e & 1] is None or halo[(q, p)][e & 1][1] == e - 2
                assert ack[(p, q)] >= e - 2
                halo[(q, p)][e & 1] = (p, e)
                ready[(q, p)] = e
                if pc[q][0] <= epochs:
                    max_lead = max(max_lead, e - pc[q][0])
            pc[p][1] = 2
        elif ph == 2:
            pc[p][1] = 3
        else:
            for s in sources[p]:
                assert halo[(p, s)][e & 1] == (s, e), f"rank {p} epoch {e} read {halo[(p, s)]} from {s}"
                ack[(s, p)] = e
            pc[p] = [e + 1, 0]
        steps += 1
        assert steps < 10_000_000
    return steps, max_lead


@pytest.mark.parametrize("seed", range(20))
def test_protocol_random_graphs_and_schedules(seed):
    rng = random.Random(seed)
    P = rng.randint(2, 8)
    if seed % 3 == 0:  # slab partition: symmetric nearest neighbours
        sends = {p: [q for q in (p - 1, p + 1) if 0 <= q < P] for p in range(P)}
    elif seed % 3 == 1:  # all-to-all (the random matrix at 8 GPUs)
        sends = {p: [q for q in range(P) if q != p] for p in range(P)}
    else:  # arbitrary, possibly asymmetric, possibly with self messages and isolated ranks
        sends = {p: [q for q in range(P) if rng.random() < 0.4] for p in range(P)}
    run(P, sends, epochs=25, rng=rng)


def test_protocol_adversarial_schedule_one_rank_far_ahead():
    """one rank runs as far ahead as the protocol allows: it must stall at the back-pressure wait, never overwrite"""
    class Greedy(random.Random):
        def choice(self, seq):
            return min(seq)  # always advance the lowest runnable rank

    sends = {0: [1], 1: [0, 2], 2: [1]}
    _, lead = run(3, sends, epochs=10, rng=Greedy(0))
    # rank 0 already writes the halo of apply e + 1 while rank 1 is still inside apply e (with ONE buffer it had to wait for
    # rank 1's boundary rows of apply e); it cannot get further ahead because its own boundary rows need rank 1's pack
    assert lead == 1


# ---------------------------------------------------------------------------------------------------------------
# CTA-level model of ONE rank's fused kernel (csrc/jp_csr.cu fused_halo_body): does forward progress depend on the order
# in which the hardware dispatches CTAs?  A rank has `slots` resident CTA slots; CTAs are dispatched in an ARBITRARY order
# (the scheduler below may start the boundary CTAs first); a CTA that waits for a flag keeps its slot.
#   claims = True   the shipped kernel: pack work is cut into items claimed from a ticket counter by the pack CTAs and by
#                   every boundary CTA before it starts to wait
#   claims = False  round 1's kernel: pack CTA i does pack item i, boundary CTAs just wait
# ---------------------------------------------------------------------------------------------------------------
def run_cta_model(P, sends, epochs, n_pack, n_int, n_bnd, slots, order, rng, claims):
    sources = {q: [p for p in range(P) if q in sends[p]] for q in range(P)}
    ready = {(q, p): 0 for p in range(P) for q in sends[p]}
    ack = {(p, q): 0 for p in range(P) for q in sends[p]}
    halo = {(q, p): [None, None] for p in range(P) for q in sends[p]}

    class Rank:
        pass

    ranks = []
    for p in range(P):
        r = Rank()
        r.epoch, r.pending, r.resident = 1, None, []
        ranks.append(r)

    def start_kernel(p):
        r = ranks[p]
        ctas = [("pack", i) for i in range(n_pack)] + [("int", i) for i in range(n_int)] + [("bnd", i) for i in range(n_bnd)]
        r.pending = order(ctas, rng)       # the dispatch order is the scheduler's choice, not the program's
        r.claimed, r.items_done, r.bnd_done = 0, 0, 0

    for p in range(P):
        start_kernel(p)
    guard = 0
    while any(r.epoch <= epochs for r in ranks):
        progressed = False
        for p in rng.sample(range(P), P):
            r = ranks[p]
            if r.epoch > epochs:
                continue
            e = r.epoch
            while r.pending and len(r.resident) < slots:   # dispatch into free slots
                kind, i = r.pending.pop(0)
                r.resident.append({"kind": kind, "i": i, "state": "claim" if kind in ("pack", "bnd") else "work", "item": None})
                progressed = True
            for cta in list(r.resident):
                if cta["state"] == "claim":
                    if claims:
                        if r.claimed < n_pack:
                            cta["item"], r.claimed = r.claimed, r.claimed + 1
                            cta["state"] = "pack_wait"
                        else:
                            cta["state"] = "work" if cta["kind"] == "bnd" else "done"
                    else:
                        cta["item"] = cta["i"] if cta["kind"] == "pack" else None
                        cta["state"] = "pack_wait" if cta["kind"] == "pack" else "work"
                    progressed = True
                elif cta["state"] == "pack_wait":
                    if all(ack[(p, q)] >= e - 2 for q in sends[p]):      # back-pressure of the two-buffer protocol
                        r.items_done += 1
                        if r.items_done == n_pack:                       # whoever finishes the last item publishes
                            for q in sends[p]:
                                assert halo[(q, p)][e & 1] is None or halo[(q, p)][e & 1][1] == e - 2
                                halo[(q, p)][e & 1] = (p, e)
                                ready[(q, p)] = e
                        cta["state"] = "claim" if claims else "done"
                        progressed = True
                elif cta["state"] == "work":
                    if cta["kind"] == "bnd":
                        if not all(ready[(p, s)] >= e for s in sources[p]):
                            continue                                     # spins, keeps its slot
                        for s in sources[p]:
                            assert halo[(p, s)][e & 1] == (s, e)
                        r.bnd_done += 1
                        if r.bnd_done == n_bnd:
                            for s in sources[p]:
                                ack[(s, p)] = e
                    cta["state"] = "done"
                    progressed = True
                if cta["state"] == "done":
                    r.resident.remove(cta)
                    progressed = True
            if not r.pending and not r.resident:                         # kernel finished: next apply
                if n_bnd == 0:
                    for s in sources[p]:
                        ack[(s, p)] = e
                r.epoch += 1
                if r.epoch <= epochs:
                    start_kernel(p)
                progressed = True
        if not progressed:
            return False                                                 # every resident CTA spins, nothing can be dispatched
        guard += 1
        assert guard < 1_000_000
    return True


def _boundary_first(ctas, rng):
    return sorted(ctas, key=lambda c: {"bnd": 0, "int": 1, "pack": 2}[c[0]])


def _shuffled(ctas, rng):
    ctas = list(ctas)
    rng.shuffle(ctas)
    return ctas


def test_fused_kernel_progress_does_not_depend_on_cta_dispatch_order():
    sends = {0: [1], 1: [0, 2], 2: [1]}
    for seed in range(10):
        rng = random.Random(seed)
        for order in (_boundary_first, _shuffled):
            assert run_cta_model(3, sends, epochs=6, n_pack=3, n_int=5, n_bnd=4, slots=2, order=order, rng=rng, claims=True)
    # the same adversarial dispatch order dead-locks round 1's kernel: every slot is held by a boundary CTA that waits for
    # a neighbour whose pack CTAs cannot be dispatched for the same reason
    assert not run_cta_model(3, sends, epochs=6, n_pack=3, n_int=5, n_bnd=4, slots=2, order=_boundary_first, rng=random.Random(0), claims=False)
