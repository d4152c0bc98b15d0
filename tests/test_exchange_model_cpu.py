"""CPU model of the in-kernel loss exchange's PROTOCOL (pytorch_retinanet_b200/csrc/loss.cu: peer_exchange).

Every rank, every step: bump the local sequence number, store (seq, value) words into slot[seq & 1][rank] of EVERY peer's
receive buffer, then spin on its own buffer's slot[seq & 1][r] for all r until the stored sequence number equals seq, and add
the values in rank order.  The claim in DESIGN.md §8 is that TWO slot parities suffice although nothing else orders the
ranks: a rank can be at most one step ahead of the slowest reader of its previous values.  Threads with random delays
play the ranks here; a (seq, value) tuple assignment stands for the single-copy-atomic 8-byte store.  Specification test
of the protocol, not of the CUDA code (tests/test_gpu_multi.py stresses that on 2 and 8 GPUs)."""
import random
import threading
import time

import pytest


def _run(world, steps, seed, jitter):
    rng = random.Random(seed)
    delays = [[rng.random() * jitter for _ in range(steps)] for _ in range(world)]
    slots = [[[(0, 0.0)] * world for _ in range(2)] for _ in range(world)]      # slots[owner][parity][sender]
    sums = [[None] * steps for _ in range(world)]
    errors = []

    def rank_main(r):
        for s in range(steps):
            time.sleep(delays[r][s])
            seq = s + 1
            par = seq & 1
            val = float((r + 1) * 1000 + s)
            for peer in range(world):
                slots[peer][par][r] = (seq, val)                # one atomic word per value
            got = []
            deadline = time.time() + 20
            for sender in range(world):
                while True:
                    q, v = slots[r][par][sender]
                    if q == seq:
                        got.append(v)
                        break
                    if q > seq or time.time() > deadline:       # overwritten before it was read, or a lost store
                        errors.append((r, s, sender, q))
                        return
                    time.sleep(0)
            sums[r][s] = sum(got)

    threads = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(60)
    return sums, errors


@pytest.mark.parametrize("world,jitter", [(2, 0.002), (4, 0.001), (8, 0.0005), (3, 0.0)])
def test_two_parities_suffice(world, jitter):
    steps = 60
    sums, errors = _run(world, steps, seed=world, jitter=jitter)
    assert not errors, errors[:3]
    for s in range(steps):
        want = sum(float((r + 1) * 1000 + s) for r in range(world))
        assert all(sums[r][s] == want for r in range(world)), (s, [sums[r][s] for r in range(world)], want)
