"""CPU: the C-ABI library loads, exports every symbol include/b200fft.h declares, the header-only vkFFT.h shim
compiles as C and C++ with the reference's struct layout, and host-side error behaviour matches the reference.
(No compute calls here: there is no GPU in the -m "not gpu" environment.)"""
import os
import re
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(built_lib):
    from vkfft_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "b200fft.h")).read()
    declared = set(re.findall(r"\b(b200fft_[a-z_]+)\s*\(", hdr))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for s in declared:
        assert hasattr(built_lib, s), s
    assert built_lib.b200fft_kernel_count() > 100
    assert built_lib.b200fft_error_string(3002) == b"VKFFT_ERROR_UNSUPPORTED_FFT_LENGTH"


def test_desc_struct_layout_matches_header(built_lib):
    import ctypes
    from vkfft_b200 import _lib
    src = r'''
    #include "b200fft.h"
    #include <stdio.h>
    #include <stddef.h>
    int main(void){ printf("%zu %zu %zu %zu %zu\n", sizeof(b200fft_desc), offsetof(b200fft_desc, buffer_stride),
        offsetof(b200fft_desc, device), sizeof(b200fft_buffers), sizeof(b200fft_plan_info)); return 0; }'''
    with tempfile.TemporaryDirectory() as td:
        open(os.path.join(td, "t.c"), "w").write(src)
        subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), os.path.join(td, "t.c"), "-o",
                               os.path.join(td, "t")])
        vals = list(map(int, subprocess.check_output([os.path.join(td, "t")]).split()))
    assert vals == [ctypes.sizeof(_lib.b200fft_desc), _lib.b200fft_desc.buffer_stride.offset,
                    _lib.b200fft_desc.device.offset, ctypes.sizeof(_lib.b200fft_buffers),
                    ctypes.sizeof(_lib.b200fft_plan_info)]


@pytest.mark.parametrize("lang", ["c", "c++"])
def test_vkfft_shim_header_layout(lang):
    """sizeof/offsetof of the drop-in structs == the reference's VKFFT_BACKEND==1 build (SURVEY.md section 7:
    VkFFTConfiguration 1168 B, VkFFTLaunchParams 80 B, buffer@152, numberBatches@272, doublePrecision@360,
    performR2C@408; measured from the reference headers)."""
    src = r'''
    #include "vkFFT.h"
    #include <stdio.h>
    #include <stddef.h>
    int main(void){ printf("%zu %zu %zu %zu %zu %zu\n", sizeof(VkFFTConfiguration), sizeof(VkFFTLaunchParams),
        offsetof(VkFFTConfiguration, buffer), offsetof(VkFFTConfiguration, numberBatches),
        offsetof(VkFFTConfiguration, doublePrecision), offsetof(VkFFTConfiguration, performR2C));
        VkFFTApplication app = VKFFT_ZERO_INIT; VkFFTConfiguration cfg = VKFFT_ZERO_INIT;
        (void)app; (void)cfg; return VkFFTGetVersion() == 10304 ? 0 : 1; }'''
    cuda = "/usr/local/cuda"
    if not os.path.exists(os.path.join(cuda, "include", "cuda.h")):
        pytest.skip("CUDA headers not present")
    with tempfile.TemporaryDirectory() as td:
        ext = "c" if lang == "c" else "cpp"
        f = os.path.join(td, "t." + ext)
        open(f, "w").write(src)
        cc = ["gcc", "-std=c99"] if lang == "c" else ["g++", "-std=c++11"]
        subprocess.check_call(cc + ["-c", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(cuda, "include"), f,
                                    "-o", os.path.join(td, "t.o")])


def test_python_api_host_side_errors(built_lib):
    import vkfft_b200 as vk
    app = vk.VkFFTApplication()
    assert vk.initializeVkFFT(None, vk.VkFFTConfiguration()) == vk.VKFFT_ERROR_EMPTY_app
    assert vk.initializeVkFFT(app, vk.VkFFTConfiguration(FFTdim=1, size=[8])) == vk.VKFFT_ERROR_INVALID_DEVICE
    assert vk.initializeVkFFT(app, vk.VkFFTConfiguration(FFTdim=0, size=[8], device=0)) == vk.VKFFT_ERROR_EMPTY_FFTdim
    assert vk.initializeVkFFT(app, vk.VkFFTConfiguration(FFTdim=5, size=[8], device=0)) == vk.VKFFT_ERROR_FFTdim_GT_MAX_FFT_DIMENSIONS
    assert vk.initializeVkFFT(app, vk.VkFFTConfiguration(FFTdim=1, size=[0], device=0)) == vk.VKFFT_ERROR_EMPTY_size
    assert vk.VkFFTAppend(app, -1, None) == vk.VKFFT_ERROR_PLAN_NOT_INITIALIZED
    assert vk.VkFFTGetVersion() == 10304
    # no GPU here: plan creation must fail loudly, never fall back to a CPU path
    import torch
    if not torch.cuda.is_available():
        assert vk.initializeVkFFT(app, vk.VkFFTConfiguration(FFTdim=1, size=[8], device=0)) == vk.VKFFT_ERROR_INVALID_DEVICE


def _build_sample(td):
    cuda = "/usr/local/cuda"
    exe = os.path.join(td, "drop_in_sample")
    subprocess.check_call(["g++", "-std=c++11", "-O1", os.path.join(ROOT, "tests", "cpp", "drop_in_sample.cpp"), "-o", exe,
                           "-I", os.path.join(ROOT, "include"), "-I", os.path.join(cuda, "include"),
                           "-L", os.path.join(ROOT, "vkfft_b200", "lib"), "-lb200fft",
                           "-L", os.path.join(cuda, "lib64"), "-L", os.path.join(cuda, "lib64", "stubs"), "-lcuda", "-lcudart",
                           "-Wl,-rpath," + os.path.join(ROOT, "vkfft_b200", "lib"), "-Wl,-rpath," + os.path.join(cuda, "lib64")])
    return exe


def test_reference_style_cpp_program_builds_against_the_shim(built_lib):
    if not os.path.exists("/usr/local/cuda/include/cuda.h"):
        pytest.skip("CUDA headers not present")
    with tempfile.TemporaryDirectory() as td:
        assert os.path.exists(_build_sample(td))


@pytest.mark.gpu
def test_reference_style_cpp_program_runs(built_lib):
    with tempfile.TemporaryDirectory() as td:
        out = subprocess.run([_build_sample(td)], capture_output=True, text=True)
        assert out.returncode == 0, out.stdout + out.stderr


def test_reference_testsuite_sources_build_against_the_shim(built_lib):
    """drop-in at source level: the reference's own VkFFT_TestSuite.cpp with its benchmark / convolution samples and utilities
    (unmodified, where they lie) compiles against include/vkFFT.h and links to libb200fft.so"""
    if not os.path.isdir("/root/reference/benchmark_scripts"):
        pytest.skip("reference tree not present")
    exe = os.path.join(ROOT, "oracle", "_ref", "VkFFT_TestSuite_b200")
    if os.path.exists(exe):
        os.unlink(exe)
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "testsuite"], stdout=subprocess.DEVNULL)
    assert os.path.exists(exe)
    syms = subprocess.run(["nm", "-D", "--undefined-only", exe], capture_output=True, text=True).stdout
    for s in ("b200fft_plan_create", "b200fft_exec", "b200fft_plan_destroy", "b200fft_plan_axis_uploads"):
        assert s in syms, s                      # the samples' initializeVkFFT / VkFFTAppend / deleteVkFFT end in the C ABI
    assert "nvrtcCompileProgram" not in syms     # nothing is JIT-compiled any more
