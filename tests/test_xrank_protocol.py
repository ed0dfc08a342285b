"""CPU: the mailbox protocol of the strict multi-GPU exchange (csrc/frontend.cu, fe_xrank_kernel), as a randomly scheduled
state machine.  Each rank, for exchange e = 1, 2, ...: stores (value, e) into slot [e & 1][me] of EVERY rank's mailbox (one atomic
8-byte word per store), then polls its own mailbox until every slot [e & 1][r] carries e, and only then goes on to e + 1.

Checked here, over random interleavings of the ranks' individual stores and polls:
* every rank reads, for every exchange, exactly the values the ranks contributed to THAT exchange (two parity slots suffice: a rank
  can run at most one exchange ahead of a peer that is still reading);
* with a single slot per rank the same schedule generator does produce wrong reads (the test has teeth).
"""
import random


def simulate(world, exchanges, slots, seed):
    rng = random.Random(seed)
    mailbox = [[[(None, 0)] * world for _ in range(slots)] for _ in range(world)]  # mailbox[rank][parity][src] = (value, epoch)
    value = lambda r, e: 1000 * e + r  # noqa: E731
    # per-rank program counter: (epoch, phase, index); phase 0 = storing to peer `index`, phase 1 = polling slot `index`
    pc = [(1, 0, 0) for _ in range(world)]
    got = [dict() for _ in range(world)]
    wrong = 0
    live = list(range(world))
    while live:
        r = rng.choice(live)
        e, phase, i = pc[r]
        par = e % slots
        if phase == 0:
            mailbox[i][par][r] = (value(r, e), e)
            pc[r] = (e, 0, i + 1) if i + 1 < world else (e, 1, 0)
        else:
            v, ep = mailbox[r][par][i]
            if ep >= e if slots == 1 else ep == e:  # the kernel compares for equality; a single slot can only ask "at least e"
                if v != value(i, e):
                    wrong += 1
                got[r].setdefault(e, []).append(v)
                if i + 1 < world:
                    pc[r] = (e, 1, i + 1)
                elif e < exchanges:
                    pc[r] = (e + 1, 0, 0)
                else:
                    live.remove(r)
    return wrong, got


def test_two_parity_slots_never_mix_exchanges():
    for world in (2, 3, 4, 8):
        for seed in range(40):
            wrong, got = simulate(world, exchanges=12, slots=2, seed=seed)
            assert wrong == 0
            for r in range(world):
                assert all(got[r][e] == [1000 * e + s for s in range(world)] for e in range(1, 13))


def test_one_slot_is_not_enough():
    # a fast rank overwrites its word for exchange e + 1 before a slow peer has read the one of exchange e
    assert sum(simulate(world, exchanges=12, slots=1, seed=seed)[0] for world in (2, 4) for seed in range(40)) > 0
