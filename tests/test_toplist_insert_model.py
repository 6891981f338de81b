"""CPU check of the generated sorted-insert blocks (radar_sounder_crw_b200/csrc/toplist_insert.inc): the inline-PTX text is
interpreted instruction by instruction (setp / predicated fma moves / predicate logic, fp32 arithmetic) and compared with a plain
sorted insert.  Pins the logic of the device code the tensor-path top-k runs, without a GPU."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = os.path.join(ROOT, "radar_sounder_crw_b200", "csrc", "toplist_insert.inc")
NO_VALUE = np.float32(-3.402823466e+38)
NO_ID = np.float32(16777216.0)


def _blocks():
    text = open(INC).read()
    out = {}
    for m in re.finditer(r"void (toplist_insert\w*)<(\d+)>\(.*?asm\((.*?)\n\s*:", text, re.S):
        name, kt, body = m.group(1), int(m.group(2)), m.group(3)
        out[(name, kt)] = [l.replace("\\n\\t", "").strip() for l in re.findall(r'"(.*?)"', body)]
    return out


def _imm(tok):
    return np.frombuffer(bytes.fromhex(tok[2:])[::-1], dtype=np.float32)[0]


def _run(lines, regs):
    """regs: list of np.float32 operands %0..; returns the updated list."""
    regs = list(regs)
    pred = {}

    def val(tok):
        tok = tok.strip()
        return regs[int(tok[1:])] if tok.startswith("%") else _imm(tok)

    for line in lines:
        line = line.rstrip(";").strip()
        if line in ("{", "}") or line.startswith(".reg"):
            continue
        guard = True
        if line.startswith("@"):
            g, line = line.split(" ", 1)
            guard = pred[g[1:]]
        op, rest = line.split(" ", 1)
        a = [t.strip() for t in rest.split(",")]
        if op == "setp.gt.f32":
            pred[a[0]] = bool(val(a[1]) > val(a[2]))
        elif op == "setp.eq.f32":
            pred[a[0]] = bool(val(a[1]) == val(a[2]))
        elif op == "setp.lt.and.f32":
            pred[a[0]] = bool(val(a[1]) < val(a[2])) and pred[a[3]]
        elif op == "or.pred":
            pred[a[0]] = pred[a[1]] or pred[a[2]]
        elif op == "fma.rn.f32":
            if guard:
                with np.errstate(over="ignore"):
                    regs[int(a[0][1:])] = np.float32(np.float64(val(a[1])) * np.float64(val(a[2])) + np.float64(val(a[3])))
        else:
            raise AssertionError("unknown instruction: " + line)
    return regs


def _reference(v, ids, x, xid, tie):
    items = list(zip(v, ids))
    pos = len(items)
    for s, (vs, is_) in enumerate(items):
        if x > vs or (tie and x == vs and xid < is_):
            pos = s
            break
    items.insert(pos, (x, xid))
    items = items[: len(v)]
    return [i[0] for i in items], [i[1] for i in items]


@pytest.mark.parametrize("key", sorted(_blocks().keys()))
def test_generated_insert_blocks_are_sorted_inserts(key):
    name, kt = key
    lines = _blocks()[key]
    tie = "tie" in name
    rs = np.random.RandomState(kt + (100 if tie else 0))
    for trial in range(40):
        v = [NO_VALUE] * kt
        ids = [NO_ID] * kt
        pool = rs.randn(6).astype(np.float32) if trial % 2 else rs.randn(400).astype(np.float32)     # odd trials: many exact ties
        for step in range(3 * kt):
            x = np.float32(pool[rs.randint(len(pool))])
            xid = np.float32(rs.randint(0, 1 << 20)) if tie else np.float32(step)
            if step % 7 == 6:
                x = np.float32(-np.inf)                 # the "mask empty" candidate of the pop: must be a no-op
            got = _run(lines, v + ids + [x, xid])
            rv, ri = _reference(v, ids, x, xid, tie)
            assert [float(t) for t in got[:kt]] == [float(t) for t in rv], (name, kt, trial, step)
            assert [float(t) for t in got[kt:2 * kt]] == [float(t) for t in ri], (name, kt, trial, step)
            v, ids = got[:kt], got[kt:2 * kt]
            assert all(v[s] >= v[s + 1] for s in range(kt - 1))
