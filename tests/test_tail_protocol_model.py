"""Model check (CPU, pure Python) of the tagged-slot exchange the fused PCG tail uses for its grid-wide sums
(apex_solver_b200/csrc/schur.cu: tail_publish / tail_gather): every CTA publishes its partial as two 64-bit words
{half of the double, 32-bit tag} into slot set `tag & 1`, and polls the G slots of that set until both words of every slot carry the
tag. Claims checked under random interleavings of single memory operations (a publish is TWO separate word stores, a poll reads
the two words of a slot separately - torn accesses included):
  * a CTA only ever accepts a value whose two halves belong to the same exchange and the same CTA (the per-word tags detect tearing);
  * two slot sets suffice: no CTA can overwrite a slot another CTA has still to read, because nobody leaves exchange t+1 before
    everybody has published there, i.e. has finished reading exchange t;
  * nobody gets stuck.
Negative control: with ONE slot set the same schedule generator finds lost updates (a slow CTA waits for a tag that a fast CTA has
already overwritten), so the check has teeth."""
import random

import pytest


def run(G, K, nsets, rng, max_steps=400000, slow=None):
    """Random-interleaving simulation. Returns ("ok", None) or (kind, detail)."""
    mem = {}                                   # (set, cta, word) -> (tag, half-id); half-id = (owner cta, exchange, word)
    # per-CTA program counter: exchange k (1..K), phase, and poll bookkeeping
    st = [{"k": 1, "phase": "pub0", "seen": set(), "cur": None} for _ in range(G)]
    done = 0
    for _ in range(max_steps):
        live = [c for c in range(G) if st[c]["k"] <= K]
        if not live:
            return "ok", None
        c = rng.choice(live)
        if slow is not None and c == slow and len(live) > 1 and rng.random() < 0.95:
            continue                           # one CTA runs twenty times slower than the others
        s = st[c]
        k = s["k"]
        sset = k % nsets
        if s["phase"] == "pub0":
            mem[(sset, c, 0)] = (k, (c, k, 0)); s["phase"] = "pub1"
        elif s["phase"] == "pub1":
            mem[(sset, c, 1)] = (k, (c, k, 1)); s["phase"] = "poll"; s["seen"] = set(); s["cur"] = None
        else:
            # poll one word of one not-yet-accepted slot; a slot is accepted when both words were read with tag k in ONE attempt
            if s["cur"] is None:
                todo = [b for b in range(G) if b not in s["seen"]]
                b = rng.choice(todo)
                w0 = mem.get((sset, b, 0), (0, None))
                s["cur"] = (b, w0)
            else:
                b, w0 = s["cur"]
                w1 = mem.get((sset, b, 1), (0, None))
                s["cur"] = None
                if w0[0] == k and w1[0] == k:
                    if w0[1] != (b, k, 0) or w1[1] != (b, k, 1):
                        return "corrupt", (c, k, b, w0, w1)
                    s["seen"].add(b)
                    if len(s["seen"]) == G:
                        s["k"] = k + 1; s["phase"] = "pub0"
                elif w0[0] > k or w1[0] > k:
                    return "lost", (c, k, b, w0[0], w1[0])     # the value of exchange k was overwritten before c read it
    return "stuck", [s["k"] for s in st]


@pytest.mark.parametrize("G,K", [(2, 8), (3, 8), (5, 6), (9, 5)])
def test_two_slot_sets_are_enough_and_tags_detect_tearing(G, K):
    rng = random.Random(1234 + G)
    for i in range(300):
        kind, detail = run(G, K, 2, rng, slow=(i % G if i % 2 else None))   # every other schedule with one very slow CTA
        assert kind == "ok", (kind, detail)


def test_one_slot_set_loses_updates_negative_control():
    rng = random.Random(7)
    kinds = {run(3, 6, 1, rng)[0] for _ in range(300)}
    assert "lost" in kinds and "corrupt" not in kinds
