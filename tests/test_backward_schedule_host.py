"""The choreography of csrc/backward_pipe.cu, modelled in Python and checked against the obvious per-pixel loop:
lanes own splats, at window step tm lane l works on pixel (tm - l) mod 32, takes the pixel's running state from lane l-1
(lane 31 -> lane 0 across batches), swaps its finished splat for the next batch's at step l, splats flagged `first` take
the unit's initial state instead of their neighbour's, batches may mix several units and end in null splats, and one last
window of nulls drains the pipeline.  The state update is order-sensitive (s -> s * a + b), so any pixel meeting the splats
of its unit out of list order, twice, or not at all shows up in the per-splat sums."""
import numpy as np


def reference(units):
    out = []
    for u in units:
        s = u["init"].copy()                      # per-pixel state at the unit's first position
        for a, b, mask in u["splats"]:
            acc = 0.0
            for p in range(32):
                if (mask >> p) & 1:
                    acc += s[p] * (p + 1)          # what the splat "sees" of pixel p
                    s[p] = s[p] * a + b
            out.append(acc)
    return out


def pipeline(units):
    queue = []                                     # (a, b, mask, unit index, first flag, output slot)
    for ui, u in enumerate(units):
        for k, (a, b, mask) in enumerate(u["splats"]):
            queue.append((a, b, mask, ui, k == 0, len(queue)))
    n_out = len(queue)
    null = (1.0, 0.0, 0, 0, False, None)
    while len(queue) % 32:
        queue.append(null)                         # the last batch is completed with null splats
    queue += [null] * 32                           # the drain window
    out = [None] * n_out
    cur = [null] * 32                              # the splat each lane holds
    acc = [0.0] * 32
    hand = [0.0] * 32                              # state a lane hands to its neighbour at the next step
    for w in range(len(queue) // 32):
        batch = queue[32 * w:32 * w + 32]
        for tm in range(32):
            got = [hand[(l + 31) & 31] for l in range(32)]   # the rotate: every lane reads before anyone writes
            for l in range(32):
                if tm == l:                        # this lane's splat has seen all 32 pixels
                    if cur[l][5] is not None:
                        out[cur[l][5]] = acc[l]
                    acc[l], cur[l] = 0.0, batch[l]
                a, b, mask, ui, first, _ = cur[l]
                p = (tm - l) & 31
                s = units[ui]["init"][p] if first else got[l]
                if (mask >> p) & 1:
                    acc[l] += s * (p + 1)
                    s = s * a + b
                hand[l] = s
    assert all(v is not None for v in out)
    return out


def test_rotating_pixel_pipeline_visits_every_pair_once_and_in_list_order():
    rng = np.random.default_rng(0)
    units = []
    for n in (1, 2, 40, 3, 0, 33, 64, 5, 1, 97, 31, 32):   # tiny units share batches, long ones span several windows
        units.append({"init": rng.uniform(0.5, 1.5, 32),
                      "splats": [(float(rng.uniform(0.5, 0.99)), float(rng.uniform(0.0, 0.1)), int(rng.integers(0, 1 << 32)))
                                 for _ in range(n)]})
    units = [u for u in units if u["splats"]]      # units without a selected splat never enter the queue
    want, got = reference(units), pipeline(units)
    assert len(want) == len(got) == sum(len(u["splats"]) for u in units)
    assert np.allclose(got, want, rtol=1e-12, atol=0.0)
