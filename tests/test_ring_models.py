"""Schedule models of the shared-memory rings of the round-2 kernels (CPU only).

The kernels refill a ring slot while other warps may still be a step behind; what makes that safe is the position of
the CTA barriers relative to the refill requests.  These models replay the kernels' request / wait / barrier / read
sequence for many random warp interleavings and check two invariants: (1) a slot is never overwritten while a warp
still has to read its previous content, (2) nothing is read before the request that fills it was issued (and, for
cp.async groups, waited for).  They mirror

* kf_mnmf_update (ssb_coop.cu): operand chunks of 32 inner indices in MUS = 3 CTA-wide slots, requested one chunk ahead
  on even steps BEHIND the CTA barrier of that step; per-warp Z2 tiles in ZST = 2 stages requested one step ahead behind
  the warp-level barrier;
* kc_cov_mma8 / kc_cov_mma4 (ssb_covmma.cu): X stages (XS = 3 resp. 5) and V chunks (2 resp. 3 slots per source)
  requested behind the per-step __syncthreads, phi double buffer written in phase A and read in phase B.
"""
import random

import pytest


class Ring:
    def __init__(self, slots):
        self.content = [None] * slots       # item id currently (being) written into the slot
        self.pending = [dict() for _ in range(slots)]  # slot -> {item: set(warps that still have to read it)}

    def request(self, slot, item, readers):
        old = self.content[slot]
        if old is not None:
            left = self.pending[slot].get(old, set())
            assert not left, "slot %d refilled with %r while warps %s still read %r" % (slot, item, sorted(left), old)
        self.content[slot] = item
        self.pending[slot][item] = set(readers)

    def read(self, slot, item, warp):
        assert self.content[slot] == item, "warp %d reads %r from slot %d which holds %r" % (warp, item, slot,
                                                                                            self.content[slot])
        self.pending[slot][item].discard(warp)


def _run_lockstep(n_warps, n_steps, body, barrier_every, rng, barrier_offset=0):
    """Warps advance one step at a time in random order.  A CTA barrier sits at the top of every step s with
    s % barrier_every == barrier_offset: a warp may enter such a step only when every warp has finished step s - 1, and
    between two barriers the warps drift freely."""
    done = [0] * n_warps  # steps completed per warp
    while min(done) < n_steps:
        cand = []
        for w in range(n_warps):
            s = done[w]
            if s >= n_steps:
                continue
            if s % barrier_every == barrier_offset and not all(d >= s for d in done):
                continue  # waiting at the barrier
            cand.append(w)
        w = rng.choice(cand)
        body(w, done[w])
        done[w] += 1


@pytest.mark.parametrize("nsteps", [1, 2, 3, 8, 33])
@pytest.mark.parametrize("seed", range(5))
def test_mnmf_update_operand_and_tile_rings(nsteps, seed):
    rng = random.Random(seed)
    MUW, MUS, ZST = 4, 3, 2
    nchunk = (nsteps + 1) // 2
    ops = Ring(MUS)
    tiles = [Ring(ZST) for _ in range(MUW)]
    warps = range(MUW)
    # prologue (before the loop, all warps): chunk 0 -> slot 0, tile 0 -> stage 0
    ops.request(0, ("chunk", 0), warps)
    for w in warps:
        tiles[w].request(0, ("tile", 0), [w])
    issued_op = set()

    def body(w, s):
        chunk = s >> 1
        # cp.async.wait_group 0 + barrier happen here (the lock-step driver models the barrier on even steps)
        if s + 1 < nsteps:
            tiles[w].request((s + 1) % ZST, ("tile", s + 1), [w])
        if s % 2 == 0 and chunk + 1 < nchunk and (chunk + 1) not in issued_op:
            # issued cooperatively by the threads of the CTA behind the barrier: modelled once, by the first warp past it
            ops.request((chunk + 1) % MUS, ("chunk", chunk + 1), warps)
            issued_op.add(chunk + 1)
        tiles[w].read(s % ZST, ("tile", s), w)
        ops.read(chunk % MUS, ("chunk", chunk), w)

    _run_lockstep(MUW, nsteps, body, barrier_every=2, rng=rng)


@pytest.mark.parametrize("n_src,xs,vs,n_warps", [(8, 3, 2, 16), (4, 5, 3, 8)])
@pytest.mark.parametrize("nsteps", [1, 2, 5, 32, 65])
@pytest.mark.parametrize("seed", range(3))
def test_covariance_gemm_rings(n_src, xs, vs, n_warps, nsteps, seed):
    rng = random.Random(100 + seed)
    nchunk = (nsteps + 1) // 2
    warps = range(n_warps)
    xring, phi = Ring(xs), Ring(2)
    vring = [Ring(vs) for _ in range(n_src)]
    for s in range(min(xs, nsteps)):
        xring.request(s % xs, ("x", s), warps)
    a_warps = {n: [w for w in warps if w % n_src == n and w < 2 * n_src] for n in range(n_src)}  # phase A readers of V_n
    for n in range(n_src):
        for c in range(min(vs, nchunk)):
            vring[n].request(c % vs, ("v", c), a_warps[n])
    state = {"x_next": min(xs, nsteps), "v_next": {n: min(vs, nchunk) for n in range(n_src)}, "phi_written": {}}

    def phase_a(w, s):
        if w < 2 * n_src:  # (source, frame half) roles
            n = w % n_src
            vring[n].read((s >> 1) % vs, ("v", s >> 1), w)
            key = ("phi", s)
            if key not in state["phi_written"]:
                phi.request(s % 2, key, warps)
                state["phi_written"][key] = True

    def phase_b(w, s):
        # behind the step's __syncthreads: lane 0 of warp 0 re-arms the X stage of step s - 1, lane 0 of warp n < n_src
        # the V slot of the chunk consumed in steps s - 2, s - 1
        if s >= 1:
            if w == 0 and s - 1 + xs < nsteps and state["x_next"] == s - 1 + xs:
                xring.request((s - 1) % xs, ("x", s - 1 + xs), warps)
                state["x_next"] += 1
            if w < n_src and s % 2 == 0:
                cn = (s >> 1) - 1 + vs
                if cn < nchunk and state["v_next"][w] == cn:
                    vring[w].request(cn % vs, ("v", cn), a_warps[w])
                    state["v_next"][w] += 1
        phi.read(s % 2, ("phi", s), w)
        xring.read(s % xs, ("x", s), w)

    # one CTA barrier per step, between phase A and phase B: model a step as two half-steps with a barrier before B
    def body(w, h):
        s, half = divmod(h, 2)
        (phase_a if half == 0 else phase_b)(w, s)

    _run_lockstep(n_warps, 2 * nsteps, body, barrier_every=2, rng=rng, barrier_offset=1)  # barrier before phase B


@pytest.mark.parametrize("seed", range(3))
def test_model_detects_a_refill_of_the_stage_in_use(seed):
    """The model is able to fail: re-arming the X stage of the CURRENT step (instead of the previous one) behind the
    barrier overwrites data other warps have not read yet."""
    rng = random.Random(seed)
    xs, n_warps, nsteps = 3, 16, 32
    warps = range(n_warps)
    xring = Ring(xs)
    for s in range(xs):
        xring.request(s % xs, ("x", s), warps)
    st = {"x_next": xs}

    def body(w, h):
        s, half = divmod(h, 2)
        if half == 1:
            if w == 0 and s + xs < nsteps and st["x_next"] == s + xs:
                xring.request(s % xs, ("x", s + xs), warps)
                st["x_next"] += 1
            xring.read(s % xs, ("x", s), w)

    with pytest.raises(AssertionError, match="refilled"):
        _run_lockstep(n_warps, 2 * nsteps, body, barrier_every=2, rng=rng, barrier_offset=1)
