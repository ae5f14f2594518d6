"""CPU: the C-ABI library loads, exports every symbol include/sarpost.h declares, and rejects bad
arguments without touching a device (no compute calls here)."""
import ctypes as C
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "sarpost.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sarpost_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported(sarpost):
    names = declared_functions()
    assert len(names) >= 14
    raw = C.CDLL(sarpost._lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), f"{n} declared in include/sarpost.h but not exported by libsarpost.so"
    assert set(sarpost._lib.SYMBOLS) == set(names), "ctypes binding and header disagree"


def test_no_torch_or_libcuda_link_dependency(sarpost):
    import subprocess
    out = subprocess.run(["ldd", sarpost._lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "libtorch" not in out and "libc10" not in out
    assert "libcuda.so" not in out  # the tensor-map encoder is resolved through cudaGetDriverEntryPoint


def test_version_and_workspace_arithmetic(sarpost):
    lib = sarpost._lib.lib
    assert lib.sarpost_version() == 100
    small = lib.sarpost_workspace_bytes(1, 8400, 1, 0, 300)
    big = lib.sarpost_workspace_bytes(16, 136000, 1, 0, 300)
    multi = lib.sarpost_workspace_bytes(1, 8400, 6, 1, 300)
    assert 0 < small < multi < big
    # 40 B per candidate slot (box 16 + score 4 + key 4 + 4x4 scratch) dominates
    assert big > 16 * 136000 * 40 and big < 16 * 137100 * 40 + (1 << 22)
    assert lib.sarpost_workspace_bytes(0, 8400, 1, 0, 300) < 0
    assert "workspace" in sarpost._lib.last_error()
    assert lib.sarpost_merge_workspace_bytes(10, 48, 300, 300) > 10 * 48 * 300 * 44


def test_argument_errors_do_not_need_a_gpu(sarpost):
    lib, L = sarpost._lib.lib, sarpost._lib
    p = L.NmsParams()
    p.conf_thres, p.iou_thres, p.max_det, p.max_nms, p.max_wh = 0.25, 0.7, 300, 30000, 7680.0
    one = C.c_float(0)
    ptr = C.addressof(one)
    # NULL tensors
    assert lib.sarpost_nms_decoded(None, 1, 6, 100, 2, C.byref(p), None, None, None, None, 0, None) == L.EINVAL
    # conf outside [0,1] mirrors the reference's assertion text (ops.py:217)
    p.conf_thres = 1.5
    assert lib.sarpost_nms_decoded(ptr, 1, 6, 100, 2, C.byref(p), ptr, ptr, None, ptr, 0, None) == L.EINVAL
    assert "Invalid Confidence threshold" in L.last_error()
    p.conf_thres, p.iou_thres = 0.25, -0.5
    assert lib.sarpost_nms_decoded(ptr, 1, 6, 100, 2, C.byref(p), ptr, ptr, None, ptr, 0, None) == L.EINVAL
    assert "Invalid IoU" in L.last_error()
    p.iou_thres, p.max_det = 0.7, 100000
    assert lib.sarpost_nms_decoded(ptr, 1, 6, 100, 2, C.byref(p), ptr, ptr, None, ptr, 0, None) == L.EUNSUPPORTED
    p.max_det = 300
    # channels < 4 + nc
    assert lib.sarpost_nms_decoded(ptr, 1, 5, 100, 2, C.byref(p), ptr, ptr, None, ptr, 0, None) == L.EINVAL
    # workspace too small / misaligned
    buf = (C.c_char * 1024)()
    base = (C.addressof(buf) + 255) // 256 * 256
    assert lib.sarpost_nms_decoded(ptr, 1, 6, 100, 2, C.byref(p), ptr, ptr, None, base, 256, None) == L.EWORKSPACE
    assert lib.sarpost_nms_decoded(ptr, 1, 6, 100, 2, C.byref(p), ptr, ptr, None, base + 4, 1 << 30, None) == L.EINVAL
    # head geometry
    h = L.Head()
    h.nl, h.batch, h.no, h.nc, h.reg_max = 1, 1, 70, 6, 8
    assert lib.sarpost_decode(C.byref(h), ptr, None) == L.EUNSUPPORTED  # reg_max != 16
    h.reg_max, h.no = 16, 69  # fewer channels than 4*reg_max + nc
    assert lib.sarpost_decode(C.byref(h), ptr, None) == L.EINVAL and "no 69" in L.last_error()
    assert lib.sarpost_stage_times(None) == L.EINVAL


def test_cpp_example_builds_and_links(sarpost, tmp_path):
    """examples/cpp_host_postprocess.cpp: the C ABI is usable from plain C++ (header compiles as C++, symbols link).
    Without a GPU the program reports the missing device through the error channel and exits 2."""
    import subprocess
    exe = tmp_path / "cpp_host_postprocess"
    libdir = os.path.dirname(sarpost._lib.LIB_PATH)
    cmd = ["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "cpp_host_postprocess.cpp"),
           "-L" + libdir, "-lsarpost", "-Wl,-rpath," + libdir, "-o", str(exe)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    run = subprocess.run([str(exe), "1"], capture_output=True, text=True, timeout=120)
    import torch
    if torch.cuda.is_available():
        assert run.returncode == 0 and "detections" in run.stdout, run.stdout + run.stderr
    else:
        assert run.returncode == 2 and "no usable CUDA device" in run.stderr


def test_cpp_plan_example_builds_and_links(sarpost, tmp_path):
    """examples/cpp_plan_serving_loop.cpp: the plan API (sarpost_plan_create / _run / _destroy) from plain C++ with the CUDA
    runtime — a per-frame loop on device buffers.  Without a GPU the program says so and exits 2."""
    import subprocess
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    exe = tmp_path / "cpp_plan_serving_loop"
    libdir = os.path.dirname(sarpost._lib.LIB_PATH)
    cmd = ["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(cuda, "include"),
           os.path.join(ROOT, "examples", "cpp_plan_serving_loop.cpp"), "-L" + libdir, "-lsarpost", "-L" + os.path.join(cuda, "lib64"), "-lcudart",
           "-Wl,-rpath," + libdir, "-Wl,-rpath," + os.path.join(cuda, "lib64"), "-o", str(exe)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    run = subprocess.run([str(exe), "12"], capture_output=True, text=True, timeout=120)
    import torch
    if torch.cuda.is_available():
        assert run.returncode == 0 and "detections" in run.stdout, run.stdout + run.stderr
    else:
        assert run.returncode == 2 and "no usable CUDA device" in run.stderr
