"""CPU: properties of the built SASS that the parity and 'B200-native' claims rely on."""
import os
import shutil
import subprocess

import pytest

from rock_art_radnet_b200 import _lib


def _sass():
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    funcs, cur = {}, None
    for line in out.splitlines():
        if "Function :" in line:
            cur = line.split("Function :")[1].strip()
            funcs[cur] = []
        elif cur is not None:
            funcs[cur].append(line)
    return funcs


def test_sm100a_and_no_contracted_multiply_add_in_roi_pool():
    funcs = _sass()
    pool = {k: v for k, v in funcs.items() if "roi_pool" in k}
    assert len(pool) >= 5
    for name, lines in pool.items():
        # TF computes a + (b-a)*t with separate multiply and add.  ptxas 12.9 contracts a packed multiply feeding a
        # packed add into FFMA2, so the kernels issue the PRODUCT as a packed FMA whose addend is an opaque -0.0
        # (d*t + (-0.0) == d*t exactly) and add with a separate FADD2 (see roipool.cu).  Every lerp of two lanes is
        # then FADD2 (b-a), FFMA2 (d*t - 0), FADD2 (a + m): no packed multiply may be left that could be contracted,
        # and a contraction would turn one of the FADD2 into an extra FFMA2 - so the counts must be exactly 2 : 1
        text = "\n".join(lines)
        assert "FMUL2" not in text, name
        n_add, n_fma = text.count("FADD2"), text.count("FFMA2")
        assert n_add == 2 * n_fma, (name, n_add, n_fma)
    slice8 = "\n".join(next(v for k, v in pool.items() if "slice_kernelILi8ELi14" in k))
    assert "FFMA2" in slice8 and "FADD2" in slice8          # packed f32x2 math is in use
    assert "UTMALDG.4D" in slice8                           # TMA tensor copy of the map slice (cp.async.bulk.tensor.4d)
    assert "SYNCS" in slice8                                # ... completing on an mbarrier
    assert "LDGSTS" in slice8                               # cp.async staging kept for maps wider than a TMA box
    assert "STG.E.EF.128" in slice8                         # streaming 16-byte stores
    band = "\n".join(next(v for k, v in pool.items() if "band_kernel" in k))
    assert "UTMALDG.4D" in band and "STG.E.EF.128" in band


def test_nms_uses_tma_bulk_copy_mbarrier_and_simd_minmax():
    funcs = _sass()
    hot = "\n".join(next(v for k, v in funcs.items() if "sort_nms_kernelINS_6BoxI32EjtLb1ELb1" in k))
    assert "UBLKCP" in hot                                   # cp.async.bulk (1-D TMA) key staging
    assert "SYNCS" in hot                                    # mbarrier arrive / try_wait
    assert "VIMNMX.S16x2" in hot and "VIADDMNMX.S16x2.RELU" in hot
    assert "MATCH.ANY" not in hot                            # ballot-based ranking
    assert "VOTE" in hot


def test_elf_is_sm_100a_only():
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = {ln.split(".")[-2] for ln in out.splitlines() if "ELF file" in ln}
    assert archs == {"sm_100a"}, archs


def test_cluster_form_uses_dsmem_bulk_copy_and_cluster_barriers():
    funcs = _sass()
    # <BoxI32, uint32 keys, uint16 idx, smem sort, kept in smem, kCluster = true, ...>
    cluster = [v for k, v in funcs.items() if "sort_nms_kernelINS_6BoxI32EjtLb1ELb1ELb1" in k]
    one_cta = [v for k, v in funcs.items() if "sort_nms_kernelINS_6BoxI32EjtLb1ELb1ELb0" in k]
    assert len(cluster) == 1 and len(one_cta) == 1
    text = "\n".join(cluster[0])
    assert "UBLKCP.S.S" in text                              # cp.async.bulk shared::cta -> shared::cluster (matrix blocks)
    assert "UBLKCP.S.G" in text                              # cp.async.bulk global -> shared (key staging)
    assert "UCGABAR_ARV" in text and "UCGABAR_WAIT" in text  # split barrier.cluster phases
    assert "UCGABAR" not in "\n".join(one_cta[0])            # the one-CTA form has no cluster traffic
    # the staged-only layout (12 anchors, 600x800 px) has a cluster form too
    assert any("sort_nms_kernelINS_6BoxI32EjjLb0ELb1ELb1ELb1" in k for k in funcs)


def test_detection_kernels_keep_float64_divide_and_float32_rules():
    funcs = _sass()
    det = {k: "\n".join(v) for k, v in funcs.items() if "class_nms_kernel" in k or "cluster_kernel" in k}
    assert len(det) == 4                                     # decode, head+NMS, records+NMS, cluster (final_nms)
    for name, text in det.items():
        assert "DFMA" in text or "MUFU.RCP64H" in text, name  # float64 decode / exact inter/(union+1e-6) divide
        if "class_nms_kernelILi0" not in name:                 # every form but decode-only builds ballot bit masks
            assert "VOTE" in text, name
