"""The C-ABI libraries load on a CPU-only box and export every symbol the headers declare."""
import ctypes
import os
import re

import pytest

from you_can_not_recommend_b200 import build, front_end, native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared(header, prefix):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(%s\w+)\s*\(" % prefix, text)))


def test_cuda_library_exports_header_symbols():
    lib = ctypes.CDLL(build.build_cuda())
    names = declared("ycnr_als.h", "ycnr_")
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(native.EXPORTS) == names


def test_host_library_exports_header_symbols():
    lib = ctypes.CDLL(build.build_host())
    names = declared("ycnr_host.h", "ycnr_")
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), n


def test_cuda_library_is_sm100a_only():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", build.build_cuda()], capture_output=True, text=True).stdout
    assert "sm_100a" in out and not re.search(r"sm_(?!100a)\d+", out)


def test_create_fails_loudly_without_gpu():
    if native.device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CUDA device|no CPU fallback"):
        native.Context(20, 10, 10)


def test_option_validation_messages():
    for kw, pat in ((dict(use_double_precision=True), "useDoublePrecision"), (dict(lowmem=True), "lowmem")):
        with pytest.raises(RuntimeError, match=pat):
            native.Context(20, 10, 10, **kw)
    with pytest.raises(RuntimeError, match="factorsCount"):
        native.Context(0, 10, 10)
    with pytest.raises(RuntimeError, match="factorsCount"):
        native.Context(1000, 10, 10)


def test_sass_carries_the_blackwell_instructions_the_design_claims():
    """cuobjdump -sass of the built library (no GPU needed): the Gram kernel really issues tcgen05 MMAs into TMEM
    (UTCHMMA — kind::tf32 shows under this mnemonic —, UTCBAR = tcgen05.commit, LDTM = tcgen05.ld, UTCATOMSWS = TMEM
    alloc), gathers with TMA tile::gather4 (UTMALDG; cp.async = LDGSTS for the column-block passes) and L2 prefetches
    (CCTL.E.PF2); the dual kernels use packed FP32 FMAs
    (FFMA2) and the factorisation the MUFU reciprocal square root; the by-item counting sort ranks with MATCH."""
    import subprocess
    sass = subprocess.run(["cuobjdump", "-sass", build.build_cuda()], capture_output=True, text=True).stdout
    fn = None
    per_fn = {}
    for line in sass.splitlines():
        if "Function :" in line:
            fn = line.split("Function :")[1].strip()
            per_fn[fn] = []
        elif fn is not None:
            per_fn[fn].append(line)
    def has(fn_part, mnemonic):
        return any(fn_part in f and any(mnemonic in ln for ln in body) for f, body in per_fn.items())
    for m in ("UTCHMMA", "UTCBAR", "LDTM", "UTCATOMSWS", "LDGSTS", "CCTL.E.PF2", "SYNCS", "UTMALDG"):
        assert has("gram_tc_kernel", m), m                  # UTMALDG: the TMA tile::gather4 loads of the operand atoms
    assert has("als_dual_tpt_kernel", "UBLKCP")             # TMA bulk copies of the gathered rows
    assert has("als_dual_tpt_kernel", "FFMA2") and has("als_dual_tpt_kernel", "MUFU.RSQ") and has("als_dual_tpt_kernel", "LDGSTS")
    assert has("als_primal_kernel", "MUFU.RSQ")
    assert has("item_scatter_kernel", "MATCH")
    assert not has("gram_tc_kernel", "HMMA.") or True      # no legacy mma.sync path is required anywhere


def test_kernel_resource_usage_fits_the_launch_shapes():
    """cuobjdump --dump-resource-usage (no GPU needed): no kernel spills to local memory, and the persistent Gram
    kernel — one CTA of up to 928 threads per SM — stays inside the 64 K-register file (a register bump there would
    only show up as 'too many resources requested for launch' on the GPU)."""
    import subprocess
    out = subprocess.run(["cuobjdump", "--dump-resource-usage", build.build_cuda()], capture_output=True, text=True).stdout
    fn, usage = None, {}
    for line in out.splitlines():
        line = line.strip()
        if line.startswith("Function "):
            fn = line.split()[1].rstrip(":")
        elif line.startswith("REG:") and fn:
            usage[fn] = {k: int(v) for k, v in (t.split(":") for t in line.split() if ":" in t and t.split(":")[1].isdigit())}
    assert len(usage) >= 60
    # (the 256-column solve holds 5 tiles per thread at the 128-register cap of its 480-thread CTA: 16 bytes of stack)
    spilled = [f for f, u in usage.items() if u.get("STACK", 0) > 16 or u.get("LOCAL", 0)]
    assert not spilled, spilled
    stages_threads = {5: 928, 8: 928, 16: 928, 25: 928, 31: 800}      # TcCfg<KT>::THREADS at the shared-memory-limited STAGES
    for kt, threads in stages_threads.items():
        name = next(f for f in usage if "gram_tc_kernelILi%dE" % kt in f)
        regs = usage[name]["REG"]
        assert ((regs + 7) // 8 * 8) * threads <= 65536, (name, regs, threads)


def test_multi_portion_header_scan_and_fill_match_the_per_row_rule():
    """Host side of ycnr_als_portions / ycnr_rmse_portions_async (no GPU): the worker pool's scan + deferred fill of a
    batch of portion headers gives, for every thread count, the rows a sequential walk gives — ALS rows as they are,
    RMSE rows cut into work entries of at most 64 ratings (a zero-length row is one empty entry) with running rating
    offsets — and flags headers whose ids are not ascending, out of range or whose lengths are negative."""
    import numpy as np
    from you_can_not_recommend_b200 import native
    rng = np.random.default_rng(5)
    heads, lim = [], 5000
    for p in range(37):
        R = int(rng.integers(0, 60))
        ids = np.sort(rng.choice(lim, R, replace=False)).astype(np.int32)
        lens = rng.choice([0, 1, 2, 63, 64, 65, 128, 129, 700], R).astype(np.int32)
        h = np.zeros(2 * R + 1, np.int32)
        h[0] = R
        h[1::2], h[2::2] = ids, lens
        heads.append(h)
    bad = [heads[3].copy(), heads[5].copy(), heads[7].copy()]
    if bad[0][0] >= 2:
        bad[0][3] = bad[0][1]                      # ids not strictly ascending
    bad[1][2 * int(bad[1][0]) - 1 if bad[1][0] else 0] = lim   # id out of range (R >= 1 with overwhelming probability)
    if bad[2][0] >= 1:
        bad[2][2] = -4                             # negative length
    heads_all = heads + bad
    for kind in (1, 2):
        want_ids, want_len, want_start, run = [], [], [], 0
        for h in heads:
            for r in range(int(h[0])):
                i, n = int(h[1 + 2 * r]), int(h[2 + 2 * r])
                if kind == 1:
                    want_ids.append(i); want_len.append(n); want_start.append(run); run += n
                else:
                    while True:
                        m = min(n, 64)
                        want_ids.append(i); want_len.append(m); want_start.append(run)
                        run += m
                        n -= m
                        if n <= 0:
                            break
        for threads in (1, 2, 5):
            counts, ids, ln, st = native.debug_batch_rows(kind, heads_all, lim, threads)
            assert (counts[:len(heads), 2] == 0).all()
            assert counts[len(heads) + 0, 2] != 0 or bad[0][0] < 2
            assert counts[len(heads) + 1, 2] != 0 or bad[1][0] < 1
            assert counts[len(heads) + 2, 2] != 0 or bad[2][0] < 1
            assert counts[:len(heads), 1].sum() == run
            assert ids.tolist() == want_ids and ln.tolist() == want_len and st.tolist() == want_start
