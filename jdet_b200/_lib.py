"""Loader / builder for libjdet_b200.so — the C-ABI library declared in include/jdet_b200.h.

The library is built IN-TREE (jdet_b200/_C/libjdet_b200.so) with nvcc for sm_100a, so it travels
with the repo snapshot to the GPU box.  There is no CPU fallback: if the library cannot be loaded,
every op raises.
"""
import ctypes
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_C")
SO = os.path.join(OUT_DIR, "libjdet_b200.so")
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]

c_f32p = ctypes.c_void_p   # raw device addresses (tensor.data_ptr())
_sz, _i, _f, _p = ctypes.c_size_t, ctypes.c_int, ctypes.c_float, ctypes.c_void_p

SIGNATURES = {
    "jdet_version": (ctypes.c_char_p, []),
    "jdet_box_iou_rotated_workspace_bytes": (_sz, [_i, _i]),
    "jdet_box_iou_rotated": (_i, [_p, _i, _p, _i, _p, _i, _p, _sz, _p]),
    "jdet_box_iou_rotated_ex": (_i, [_p, _i, _p, _i, _p, _i, _i, _p, _sz, _p]),
    "jdet_nms_rotated_workspace_bytes": (_sz, [_i, _i]),
    "jdet_nms_rotated": (_i, [_p, _i, _i, _p, _f, _p, _p, _sz, _p]),
    "jdet_nms_rotated_ex": (_i, [_p, _i, _i, _p, _f, _i, _p, _p, _sz, _p]),
    "jdet_argsort_desc_workspace_bytes": (_sz, [_i]),
    "jdet_argsort_desc": (_i, [_p, _i, _p, _p, _sz, _p]),
    "jdet_nms_poly_workspace_bytes": (_sz, [_i]),
    "jdet_nms_poly": (_i, [_p, _i, _p, ctypes.c_double, _i, _p, _p, _sz, _p]),
    "jdet_pack_detections_range": (_i, [_p, _i, _p, _p, _p, _f, _f, _i, _p, _p]),
    "jdet_pack_detections": (_i, [_p, _i, _i, _p, _p, _p, _i, _p, _p]),
    "jdet_roi_align_rotated_workspace_bytes": (_sz, [_i, _i, _i, _i, _i, _i, _i, _i]),
    "jdet_roi_align_rotated": (_i, [_i, _p, _i, _i, _i, _i, _p, _i, _i, _i, _f, _i, _p, _p, _sz, _p]),
    "jdet_roi_align_rotated_nhwc_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "jdet_roi_align_rotated_nhwc": (_i, [_i, _p, _i, _i, _i, _i, _p, _i, _i, _i, _f, _i, _p, _p, _sz, _p]),
    "jdet_roi_align_rotated_fpn_workspace_bytes": (_sz, [_i, _i, _i, _p, _p, _i, _i, _i, _i]),
    "jdet_roi_align_rotated_fpn": (_i, [_i, _p, _i, _i, _i, _p, _p, _p, _p, _i, _i, _i, _i, _f, _f, _f, _f, _f, _p, _p, _sz, _p]),
    "jdet_roi_align_rotated_backward_workspace_bytes": (_sz, [_i, _i, _i, _i, _i, _i, _i, _i]),
    "jdet_roi_align_rotated_backward": (_i, [_i, _p, _p, _i, _i, _i, _i, _i, _i, _i, _f, _i, _p, _p, _sz, _p]),
    "jdet_feature_refine": (_i, [_p, _p, _i, _i, _i, _i, _i, _f, _p, _p]),
    "jdet_feature_refine_multi": (_i, [_p, _p, _i, _i, _i, _p, _p, _p, _i, _p, _p]),
    "jdet_feature_refine_backward": (_i, [_p, _p, _i, _i, _i, _i, _i, _f, _p, _p]),
    "jdet_align_conv_offset": (_i, [_p, _i, _i, _i, _f, _i, _p, _p]),
    "jdet_deform_conv_forward": (_i, [_p, _p, _p] + [_i] * 16 + [_p, _p]),
    "jdet_deform_im2col": (_i, [_p, _p] + [_i] * 13 + [_p, _p]),
    "jdet_deform_col2im": (_i, [_p, _p] + [_i] * 13 + [_p, _p]),
    "jdet_deform_col2im_coord": (_i, [_p, _p, _p] + [_i] * 13 + [_p, _p]),
    "jdet_align_conv_forward_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "jdet_align_conv_forward": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _f, _p, _p, _sz, _p]),
    "jdet_align_conv_forward_multi_workspace_bytes": (_sz, [_i, _i, _i, _p, _p, _i]),
    "jdet_align_conv_forward_multi": (_i, [_p, _p, _p, _i, _i, _i, _p, _p, _i, _p, _p, _i, _p, _sz, _p]),
}


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _stale():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, jobs=None):
    """nvcc every csrc/*.cu for sm_100a and link jdet_b200/_C/libjdet_b200.so."""
    if not force and not _stale():
        return SO
    os.makedirs(OUT_DIR, exist_ok=True)
    objs, procs = [], []
    for src in sources():
        obj = os.path.join(OUT_DIR, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out.decode(errors="replace"))
        if p.returncode:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    link = ["nvcc", "-shared", "-o", SO] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
    subprocess.check_call(link)
    return SO


_lib = None


def lib():
    """ctypes handle with argtypes set.  Raises (never falls back) if the library is missing."""
    global _lib
    if _lib is None:
        alt = os.environ.get("JDET_B200_LIB")      # a prebuilt library elsewhere (A/B timing of two builds: tools/ab_libs.py)
        if alt:
            if not os.path.exists(alt):
                raise RuntimeError("JDET_B200_LIB=%s does not exist" % alt)
            L = ctypes.CDLL(alt)
            for name, (res, args) in SIGNATURES.items():
                if not hasattr(L, name):       # an older build under A/B: entry points added since are simply absent
                    continue
                fn = getattr(L, name)
                fn.restype = res
                fn.argtypes = args
            _lib = L
            return _lib
        if _stale():
            try:
                build()
            except Exception as e:  # no nvcc on this box and no prebuilt library
                if not os.path.exists(SO):
                    raise RuntimeError(
                        "libjdet_b200.so is not built and could not be built (%s); "
                        "jdet_b200 has no CPU fallback — run `python -c 'import __graft_entry__ as g; g.build()'`" % e)
                # a library older than the sources may not match SIGNATURES any more: tolerated only on request
                if os.environ.get("JDET_B200_ALLOW_STALE_LIB") != "1":
                    raise RuntimeError(
                        "libjdet_b200.so is older than jdet_b200/csrc and the rebuild failed (%s); refusing to load a library "
                        "whose ABI may differ from the Python bindings (set JDET_B200_ALLOW_STALE_LIB=1 to override)" % e)
                sys.stderr.write("jdet_b200: WARNING: loading a STALE libjdet_b200.so (rebuild failed: %s)\n" % e)
        L = ctypes.CDLL(SO)
        missing = [name for name in SIGNATURES if not hasattr(L, name)]
        if missing:
            raise RuntimeError("libjdet_b200.so does not export %s: it was built from other sources than these bindings" % missing)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


class JDetError(RuntimeError):
    pass


_ERR = {-1: "bad argument", -2: "workspace too small", -3: "unsupported size/shape"}


def check(code, what):
    if code == 0:
        return
    if code < 0:
        raise JDetError("%s: %s" % (what, _ERR.get(code, "error %d" % code)))
    raise JDetError("%s: CUDA error %d" % (what, code))
