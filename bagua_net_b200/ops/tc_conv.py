"""3x3 convolution on the tcgen05 tensor cores (csrc/cuda/tc_gemm.cu, ``kConv`` instantiations).

    y  = tc_conv.conv3x3(x, w, bias, relu=True)     # NHWC bf16, stride 1, pad 1; bias + ReLU applied from TMEM
    dx = tc_conv.conv3x3_dgrad(gy, w)               # same filter, read MN-major with flipped taps (no rotated copy)

Implicit GEMM: the 128 accumulator rows of a tile are a patch of output pixels (bw x bh pixels of bn images); for every
(filter tap, 64-channel block) the patch's shifted input arrives as ONE 4-D TMA box of the NHWC activation, the padding
being the TMA unit's zero fill.  No im2col buffer exists anywhere.  Filters are used exactly as torch stores a
channels_last ``Conv2d.weight`` ([Cout][3][3][Cin]).

The layers above (``fused_nn.ConvBiasReLU``) ask ``choose()`` which implementation to run for a given shape: on first
(eager) use each candidate is timed on the actual tensors with CUDA events and the winner is cached for the process —
so a shape where cuDNN's kernel is faster keeps cuDNN.  ``BNET_TC_CONV=1`` forces the tcgen05 kernels wherever the
shape allows, ``=0`` disables them, default ``auto``.

The reference has no compute path; this is part of the B200 training engine around the transport."""
from __future__ import annotations

import ctypes as C
import os

import torch

from ..utils.native import load
from . import tc_linear

_lib = None
LAUNCHES = 0
_choice: dict = {}          # (kind, N, H, W, Cin, Cout) -> "tc" | "cudnn"
TIMINGS: dict = {}          # the same key -> {"tc": us, "cudnn": us}  (what the autotuner measured; bench.py reports it)


def _L():
    global _lib
    if _lib is None:
        _lib = load()
        vp, i = C.c_void_p, C.c_int
        _lib.bnet_tc_conv3x3.argtypes = [vp, vp, vp, vp, i, i, i, i, i, i, vp, vp]
        _lib.bnet_tc_conv3x3_dgrad.argtypes = [vp, vp, vp, i, i, i, i, i, vp, vp]
        _lib.bnet_tc_last_error.restype = C.c_char_p
    return _lib


def mode() -> str:
    v = os.environ.get("BNET_TC_CONV", "auto").lower()
    return {"1": "on", "0": "off", "on": "on", "off": "off"}.get(v, "auto")


def _nhwc(t: torch.Tensor) -> bool:
    return t.dim() == 4 and t.is_contiguous(memory_format=torch.channels_last)


def shape_ok(x: torch.Tensor, w: torch.Tensor, dgrad: bool = False) -> bool:
    """3x3 / stride 1 / pad 1 is checked by the caller; here: dtype, layout and the kernel's channel constraints."""
    if not (x.is_cuda and x.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and _nhwc(x)):
        return False
    cout, cin = w.shape[0], w.shape[1]
    if tuple(w.shape[2:]) != (3, 3) or not w.is_contiguous(memory_format=torch.channels_last):
        return False
    if w.data_ptr() % 16 or x.data_ptr() % 16:
        return False
    if dgrad:
        return cout % 64 == 0 and cin % 64 == 0
    return cin % 64 == 0 and cout % 8 == 0


def usable() -> bool:
    return mode() != "off" and tc_linear.supported() and tc_linear.trusted()


def conv3x3(x: torch.Tensor, w: torch.Tensor, bias: torch.Tensor | None = None, relu: bool = False,
            out: torch.Tensor | None = None) -> torch.Tensor:
    """relu?(conv2d(x, w, bias, stride=1, padding=1)) — x [N,Cin,H,W] channels_last, w [Cout,Cin,3,3] channels_last, bf16."""
    global LAUNCHES
    n, cin, h, wd = x.shape
    cout = w.shape[0]
    if out is None:
        out = torch.empty((n, cout, h, wd), device=x.device, dtype=torch.bfloat16, memory_format=torch.channels_last)
    L = _L()
    rc = L.bnet_tc_conv3x3(x.data_ptr(), w.data_ptr(), bias.data_ptr() if bias is not None else None, out.data_ptr(), n, h, wd, cin,
                           cout, 1 if relu else 0, tc_linear._err_flag(x.device.index).data_ptr(), tc_linear._stream())
    if rc < 0:
        raise RuntimeError(f"bnet_tc_conv3x3: {L.bnet_tc_last_error().decode()}")
    LAUNCHES += rc
    tc_linear.LAUNCHES += rc
    return out


def conv3x3_dgrad(gy: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """Input gradient of the same convolution: gy [N,Cout,H,W] channels_last -> dx [N,Cin,H,W] channels_last."""
    global LAUNCHES
    n, cout, h, wd = gy.shape
    cin = w.shape[1]
    dx = torch.empty((n, cin, h, wd), device=gy.device, dtype=torch.bfloat16, memory_format=torch.channels_last)
    L = _L()
    rc = L.bnet_tc_conv3x3_dgrad(gy.data_ptr(), w.data_ptr(), dx.data_ptr(), n, h, wd, cin, cout,
                                 tc_linear._err_flag(gy.device.index).data_ptr(), tc_linear._stream())
    if rc < 0:
        raise RuntimeError(f"bnet_tc_conv3x3_dgrad: {L.bnet_tc_last_error().decode()}")
    LAUNCHES += rc
    tc_linear.LAUNCHES += rc
    return dx


def _time_us(fn, iters: int = 5) -> float:
    for _ in range(2):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def choose(kind: str, x: torch.Tensor, w: torch.Tensor, cudnn_fn, tc_fn, check=None) -> str:
    """Which implementation runs this (kind, shape): decided once per process by timing both on the actual tensors (never
    inside a CUDA-graph capture: an undecided shape keeps cuDNN there), after comparing their results (`check(a, b)`)."""
    n, c, h, wd = x.shape
    key = (kind, n, h, wd, w.shape[1], w.shape[0])
    got = _choice.get(key)
    if got is not None:
        return got
    m = mode()
    if m == "off" or not usable() or not shape_ok(x, w, dgrad=(kind == "dgrad")):
        _choice[key] = "cudnn"
        return "cudnn"
    if torch.cuda.is_current_stream_capturing():
        return "tc" if m == "on" else "cudnn"
    try:
        a, b = tc_fn(), cudnn_fn()
        torch.cuda.synchronize()
        if tc_linear.last_error(x.device.index):
            raise RuntimeError("pipeline watchdog")
        if check is not None and not check(a, b):
            raise RuntimeError("results differ from cuDNN")
        del a, b
        if m == "on":
            _choice[key] = "tc"
            return "tc"
        t_tc, t_cudnn = _time_us(tc_fn), _time_us(cudnn_fn)
        TIMINGS[key] = {"tc": round(t_tc, 1), "cudnn": round(t_cudnn, 1)}
        _choice[key] = "tc" if t_tc <= t_cudnn else "cudnn"
    except Exception as ex:   # noqa: BLE001 — anything unexpected keeps the library path for this shape
        TIMINGS[key] = {"error": f"{type(ex).__name__}: {str(ex)[:80]}"}
        _choice[key] = "cudnn"
    return _choice[key]


def close(a: torch.Tensor, b: torch.Tensor) -> bool:
    d = (a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-6)
    return bool(d == d and d < 2e-2)


def self_check(device=None, verbose: bool = False) -> bool:
    """Forward and input-gradient kernels against cuDNN in bf16: exact patches, ragged ones, several images per patch."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
    g = torch.Generator(device=dev).manual_seed(99)
    ok = True
    for n, cin, cout, hw in ((2, 64, 64, 32), (3, 64, 128, 14), (2, 128, 256, 28), (5, 256, 64, 7), (1, 64, 320, 20)):
        x = torch.randn(n, cin, hw, hw, device=dev, generator=g).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        w = (torch.randn(cout, cin, 3, 3, device=dev, generator=g) * (1.0 / (3 * cin ** 0.5))).to(torch.bfloat16).contiguous(
            memory_format=torch.channels_last)
        b = torch.randn(cout, device=dev, generator=g).to(torch.bfloat16)
        y = conv3x3(x, w, b, relu=True)
        ref = torch.relu(torch.nn.functional.conv2d(x, w, b, 1, 1))
        e1 = ((y.float() - ref.float()).norm() / ref.float().norm()).item()
        errs = [e1]
        if cout % 64 == 0:
            gy = torch.randn(n, cout, hw, hw, device=dev, generator=g).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
            dx = conv3x3_dgrad(gy, w)
            dref = torch.ops.aten.convolution_backward(gy, x, w, None, [1, 1], [1, 1], [1, 1], False, [0, 0], 1, [True, False, False])[0]
            errs.append(((dx.float() - dref.float()).norm() / dref.float().norm()).item())
        torch.cuda.synchronize()
        bad = tc_linear.last_error(dev.index)
        if verbose:
            print(f"[tc_conv.self_check] N={n} {cin}->{cout} {hw}x{hw}: relative L2 errors {errs} watchdog {bad}")
        ok = ok and bad == 0 and all(e == e and e < 2e-2 for e in errs)
    return ok
