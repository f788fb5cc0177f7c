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
