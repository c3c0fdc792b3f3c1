"""Model of the cp.async schedule of kf_cov_ip1_basis' second pass (ssspy_b200/csrc/ssb_coop.cu), forward and
backwards: the control flow of the kernel's issue / commit / wait_group<1> sequence restated step by step, checking
that (1) the V chunk and the X stage a step reads were issued in a group that wait_group<1> has retired, (2) they hold
the right chunk / frame step, (3) no copy still in flight targets a buffer that is being read.  The backwards variant
cannot be run without a GPU; this pins its schedule (incl. the odd-step-count case, where the last chunk holds a
single step and its predecessor must be requested in the prologue)."""
import pytest

XST = 3  # stages of the X ring


def simulate(nsteps, rev):
    nchunk = (nsteps + 1) // 2
    groups, cur = [], []
    content_v, content_x = {0: None, 1: None}, {0: None, 1: None, 2: None}
    applied = 0

    def commit():
        nonlocal cur
        groups.append(cur)
        cur = []

    def retire(upto):
        nonlocal applied
        for gi in range(applied, upto):
            for kind, buf, what in groups[gi]:
                (content_v if kind == "V" else content_x)[buf] = what
        applied = max(applied, upto)

    first = nsteps - 1 if rev else 0
    fetch = [first]

    def next_x():
        s = fetch[0]
        fetch[0] += -1 if rev else 1
        return s

    # prologue
    if rev:
        cur.append(("V", (first >> 1) & 1, first >> 1))
        if (first >> 1) >= 1:
            cur.append(("V", ((first >> 1) - 1) & 1, (first >> 1) - 1))
    else:
        cur.append(("V", 0, 0))
    cur.append(("X", 0, next_x()))
    commit()
    if nsteps > 1:
        cur.append(("X", 1, next_x()))
    commit()
    rd, wr = 0, 2
    for r in range(nsteps):
        s = nsteps - 1 - r if rev else r
        retire(len(groups) - 1)  # cp.async.wait_group 1
        if r + 2 < nsteps:
            cur.append(("X", wr, next_x()))
        if rev:
            if (s & 1) and r > 0 and (s >> 1) >= 1:
                cur.append(("V", ((s >> 1) - 1) & 1, (s >> 1) - 1))
        elif (s & 1) == 0 and (s >> 1) + 1 < nchunk:
            cur.append(("V", ((s >> 1) + 1) & 1, (s >> 1) + 1))
        commit()
        c = s >> 1
        assert content_v[c & 1] == c, ("V chunk not resident", nsteps, rev, r)
        assert content_x[rd] == s, ("X stage not resident", nsteps, rev, r)
        for gi in range(applied, len(groups)):
            for kind, buf, _ in groups[gi]:
                assert buf != ((c & 1) if kind == "V" else rd), ("copy in flight into a buffer being read", nsteps, rev, r)
        rd, wr = (rd + 1) % XST, (wr + 1) % XST


@pytest.mark.parametrize("rev", [False, True])
def test_second_pass_schedule(rev):
    for nsteps in range(1, 70):
        simulate(nsteps, rev)
