"""CPU checks of the drop-in boundary: the C-ABI library builds, loads, and exports every symbol the header
declares; argument validation fails loudly without a GPU (no compute calls are made here)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build_cuda()
    import deeplab_b200  # noqa: F401
    from deeplab_b200 import _lib
    return _lib


def test_header_symbols_are_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "deeplab_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(dlb_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    L = lib.lib()
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, missing
    assert declared == set(lib.EXPORTS), declared ^ set(lib.EXPORTS)
    m = re.search(r"#define\s+DLB_ABI_VERSION\s+(\d+)", open(os.path.join(ROOT, "include", "deeplab_b200.h")).read())
    assert L.dlb_version() == int(m.group(1))
    assert L.dlb_launch_count() == 0


def test_struct_sizes_match_header(lib):
    """ctypes mirrors must have the C layout (pointer/int packing): compile a tiny C probe against the header."""
    import subprocess
    import tempfile
    names = {"dlb_pw_gemm_params": lib.PwGemmParams, "dlb_pw_wgrad_params": lib.PwWgradParams,
             "dlb_dw_conv_params": lib.DwConvParams, "dlb_dw_conv_bwd_params": lib.DwConvBwdParams,
             "dlb_stem_conv_params": lib.StemConvParams, "dlb_bn_apply_params": lib.BnApplyParams,
             "dlb_bn_bwd_params": lib.BnBwdParams, "dlb_softmax_ce_params": lib.SoftmaxCeParams,
             "dlb_crf_config": lib.CrfConfig, "dlb_sepconv_fused_params": lib.SepconvFusedParams,
             "dlb_bn_fin": lib.BnFin, "dlb_aug_params": lib.AugParams}
    src = '#include <stdio.h>\n#include "deeplab_b200.h"\nint main(){' + "".join(
        f'printf("{n} %zu\\n", sizeof({n}));' for n in names) + "return 0;}"
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "p.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "p"), os.path.join(d, "p.c")])
        out = subprocess.check_output([os.path.join(d, "p")], text=True)
    for line in out.strip().splitlines():
        n, sz = line.split()
        assert C.sizeof(names[n]) == int(sz), (n, C.sizeof(names[n]), sz)


def test_invalid_arguments_fail_loudly(lib):
    L = lib.lib()
    p = lib.PwGemmParams()
    assert L.dlb_pw_gemm(C.byref(p), None) == -1
    assert b"null pointer" in L.dlb_last_error()
    cfg = lib.CrfConfig(0, 0, 0, 0, 0, 0, 0, 0, 0)
    assert L.dlb_crf_workspace_bytes(C.byref(cfg)) == 0
    assert L.dlb_device_ok() == 0          # no GPU in the build container


def test_product_has_no_cpu_fallback(lib):
    import torch
    from deeplab_b200 import ops
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    a = torch.zeros(128, 64)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.pw_gemm(a, a, a)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "keras-segmentation-deeplab-v3.1_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py") and fn != "selftest.py":
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src.replace("# oracle", ""), fn


def test_pw_gemm_plan_fits_every_layer_shape(lib):
    """The tensor-core GEMM's tiling plan (host arithmetic) must fit TMEM (512 columns) and shared memory (227 KB)
    with a >= 2-stage pipeline for every (K, N) of the MobileNetV2 / Xception graphs, forward and dgrad."""
    L = lib.lib()
    L.dlb_pw_gemm_plan.argtypes = [C.c_int] * 5 + [C.POINTER(C.c_int)]
    chans = [(16, 96), (96, 24), (24, 144), (144, 24), (144, 32), (32, 192), (192, 32), (192, 64), (64, 384), (384, 64),
             (384, 96), (96, 576), (576, 96), (576, 160), (160, 960), (960, 160), (960, 320), (32, 16), (320, 256),
             (512, 256), (256, 21), (256, 1344), (2048, 256), (1280, 256), (304, 256), (256, 48), (728, 728),
             (1536, 2048), (1024, 1536), (256, 256), (128, 256), (256, 728), (728, 1024)]
    M = 16 * 64 * 64
    F16, F32 = lib.F16, lib.F32
    for k, n in chans:
        for (kk, nn) in ((k, n), (n, k)):           # forward and backward-data
            for od, shuf in ((F16, 0), (F32, 0), (F16, 8 if nn == 1344 else 0)):
                plan = (C.c_int * 19)()
                assert L.dlb_pw_gemm_plan(M, nn, kk, od, shuf, plan) == 0, (kk, nn, od, shuf)
                sets, chunk_n, n_chunks, cpg, n_groups, acc_cols, acc_stages, alt, stages, smem = list(plan)[:10]
                grid, ctas = plan[10], list(plan)[11:]
                # one column group per CTA, every group served, grid within one wave of the 148 SMs
                assert 1 <= n_groups <= 8 and all(c >= 1 for c in ctas[:n_groups]) and not any(ctas[n_groups:])
                assert sum(ctas) == grid <= 148
                if n_groups > 1:      # group boundaries on 64-column blocks (bulk tensor store of staged tiles)
                    assert (cpg * chunk_n) % 64 == 0, (kk, nn, list(plan))
                # an epilogue warp sees at most three 64-column blocks per tile (register-resident statistics)
                per_warp = -(-acc_cols // 64) if alt else -(-acc_cols // (64 * sets))
                assert per_warp <= 3 or sets == 2, (kk, nn, list(plan))
                assert sets == (4 if (od != F32 and shuf == 0) else 2)
                assert chunk_n % 16 == 0 and 16 <= chunk_n <= 256 and chunk_n * n_chunks >= nn
                assert acc_cols * acc_stages <= 512, (kk, nn, list(plan))
                assert 2 <= stages <= 8 and smem <= 227 * 1024, (kk, nn, list(plan))
                assert acc_stages <= sets if alt else acc_stages <= 2
                if alt:
                    assert n_groups == 1
            # fp32 operands (3xTF32 instance, bit 17): 3 warp sets (2 drain + 1 splits the tiles), 32-element k-blocks,
            # hi + lo copies of both operands per stage, at least two stages
            plan = (C.c_int * 19)()
            assert L.dlb_pw_gemm_plan(M, nn, kk, F32, 1 << 17, plan) == 0, (kk, nn)
            sets, chunk_n, n_chunks, cpg, n_groups, acc_cols, acc_stages, alt, stages, smem = list(plan)[:10]
            assert sets == 3 and cpg == 1 and acc_cols <= 256 and acc_cols * acc_stages <= 512
            assert stages >= 2 and smem <= 227 * 1024 and 1 <= n_groups <= 8, (kk, nn, list(plan))


def test_every_kernel_honours_the_dependent_launch_contract():
    """Every kernel is launched through launch_k() with programmatic stream serialization, so every `__global__`
    body must execute griddepcontrol.wait (pdl_prologue / pdl_wait) before its first global access -- a kernel without
    it would race with its predecessor.  Static check over csrc/: no raw <<<>>> launches remain, and each kernel body
    contains the wait ahead of any dereference of a kernel parameter pointer we can recognise (`a.`/`g.`/`p->` loads
    are allowed before it only for scalars; the check is 'wait present and within the first statements')."""
    import glob
    csrc = os.path.join(ROOT, "keras-segmentation-deeplab-v3.1_b200", "csrc")
    n_kernels = 0
    for path in sorted(glob.glob(os.path.join(csrc, "*.cu"))):
        src = open(path).read()
        code = re.sub(r"//[^\n]*", "", src)
        assert "<<<" not in code, f"{path}: raw kernel launch bypasses launch_k()"
        pos = 0
        while True:
            i = code.find("__global__", pos)
            if i < 0:
                break
            # parameter list: first '(' whose preceding identifier is not __launch_bounds__
            k = i
            while True:
                k = code.find("(", k)
                ident = re.search(r"([A-Za-z_]\w*)\s*$", code[:k]).group(1)
                depth, e = 0, k
                while True:
                    depth += code[e] == "("
                    depth -= code[e] == ")"
                    if depth == 0:
                        break
                    e += 1
                if ident == "__launch_bounds__":
                    k = e + 1
                    continue
                break
            b = code.find("{", e)
            depth, j = 0, b
            while True:
                depth += code[j] == "{"
                depth -= code[j] == "}"
                if depth == 0:
                    break
                j += 1
            body = code[b:j]
            assert "pdl_prologue();" in body or "pdl_wait();" in body, f"{path}: kernel {ident} never waits on its predecessor"
            n_kernels += 1
            pos = j
    assert n_kernels >= 55, n_kernels
