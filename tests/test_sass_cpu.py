"""What the compiler made of the matcher (runs without a GPU: cuobjdump reads the sm_100a code of the built library).
The design rests on three code-generation facts: the inner loop runs on packed FP32 instructions,
the search area is staged by asynchronous global->shared copies, and the resident-CTA register caps do not spill
more than a handful of values."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "mimc3_b200", "libmimc3cu.so")


@pytest.fixture(scope="module")
def sass_rows():
    if shutil.which("cuobjdump") is None or not os.path.exists(LIB):
        pytest.skip("cuobjdump or the built library is not available")
    out = subprocess.run(["bash", os.path.join(ROOT, "scripts", "sass_summary.sh"), LIB], stdout=subprocess.PIPE, text=True,
                         check=True).stdout.splitlines()
    cols = out[1].split()
    assert cols[0] == "kernel"
    rows = {}
    for line in out[2:]:
        name, _, rest = line.rpartition(">") if ">" in line else (line.split()[0][:-1], "", " ".join(line.split()[1:]))
        vals = rest.split()
        key = (name + ">").strip() if ">" in line else line.split()[0]
        rows[key] = dict(zip(cols[1:], (int(v) for v in vals[-(len(cols) - 1):])))
    return rows


def test_matcher_inner_loop_is_packed_fp32(sass_rows):
    m2 = {k: v for k, v in sass_rows.items() if k.startswith("match2_kernel<")}
    assert len(m2) >= 20, sorted(sass_rows)
    for k, v in m2.items():
        assert v["FFMA2"] > 0 and v["FADD2"] > 0, (k, v)
        assert v["UTMALDG"] == 0


def test_search_area_is_staged_by_async_copies(sass_rows):
    for k in ("match2_kernel<40, 256, false, 4>", "match2_kernel<30, 128, false, 3>", "match2_kernel<7, 32, false, 4>"):
        assert sass_rows[k]["LDGSTS"] > 0, (k, sass_rows[k])
    # half-width 15 keeps its 16-byte loads
    assert sass_rows["match2_kernel<15, 32, false, 2>"]["LDG.128"] >= 20


def test_register_caps_do_not_spill(sass_rows):
    for k, v in sass_rows.items():
        if k.startswith("match2_kernel<"):
            assert v["STL"] <= 4 and v["LDL"] <= 4, (k, v)
