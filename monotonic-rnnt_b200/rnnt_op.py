"""Host-side mirror of the reference's PyTorch operator interface for the loss-and-gradient path.

Same names, argument meaning and error behaviour as pytorch_binding/monotonic_rnnt_op.py
(MonotonicRNNTFunction :19-118, monotonic_rnnt_loss :121-163, MonotonicRNNTLoss :166-217), but the
work goes through the flat C ABI of libmonotonic_rnnt.so (ctypes) instead of a JIT-built pybind11
module.  PyTorch is used for device memory, streams and autograd plumbing only.

Differences, all on the host side of the boundary:
  * CUDA tensors only.  CPU tensors raise: this product has no CPU path.
  * gradients are allocated with ``empty_like`` -- the kernels write every element exactly once
    (the reference allocates ``zeros_like``, monotonic_rnnt_op.py:32-36: one extra pass over N floats).
  * costs never visit the host in the autograd path (the reference copies them D2H and back,
    monotonic_rnnt_op.py:37,90).
  * ``MonotonicRNNTLoss`` works (the reference module passes kwargs to ``Function.apply`` and reads an
    undefined ``self.blank``, monotonic_rnnt_op.py:207-215).
"""
from __future__ import annotations

import contextlib
import ctypes
from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib


_NULL_CONTEXT = contextlib.nullcontext()


def _check_inputs(acts, labels, input_lengths, label_lengths) -> None:
    if not acts.is_cuda:
        raise RuntimeError("monotonic-rnnt_b200 is GPU-only: acts must be a CUDA tensor (no CPU fallback)")
    for name, t in (("labels", labels), ("input_lengths", input_lengths), ("label_lengths", label_lengths)):
        if not t.is_cuda:
            raise RuntimeError(f"{name} must be a CUDA tensor")  # monotonic_rnnt.cu:85-88
        if t.dtype != torch.int32:
            raise TypeError(f"{name} must be int32")             # data_ptr<int>() in monotonic_rnnt.cu:99-101
    if acts.dtype not in (torch.float32, torch.bfloat16):
        raise TypeError("acts must be float32 (the reference's type, monotonic_rnnt.cu:84) or bfloat16 (extension)")
    if acts.dim() not in (2, 4):
        raise ValueError("acts must be the packed 2-D tensor [sum_b T_b*(S_b+1), V] or the padded 4-D tensor "
                         "[B, T, U, V]")
    if acts.dim() == 4 and (acts.shape[0] != input_lengths.shape[0] or labels.dim() != 2):
        raise ValueError("padded acts [B, T, U, V] need labels [B, S] and one length per utterance")
    if not (acts.is_contiguous() and labels.is_contiguous() and input_lengths.is_contiguous()
            and label_lengths.is_contiguous()):
        raise ValueError("inputs must be contiguous")


def workspace_size(input_lengths_host, label_lengths_host, V: int) -> int:
    """Bytes of device workspace for a batch (host-only query; validates the lengths)."""
    T = np.ascontiguousarray(np.asarray(input_lengths_host, dtype=np.int32))
    S = np.ascontiguousarray(np.asarray(label_lengths_host, dtype=np.int32))
    out = ctypes.c_size_t(0)
    st = _lib.load().mrnnt_get_workspace_size(T.ctypes.data, S.ctypes.data, int(T.shape[0]), int(V), ctypes.byref(out))
    _lib.check(st, "mrnnt_get_workspace_size")
    return int(out.value)


class LossHandle:
    """One bound batch = one ``GpuRNNTWorkspaceManager<float>`` plus its workspace.

    Mirrors how the reference bindings drive the C++ classes (pytorch_binding/monotonic_rnnt.cu:99-111):
    construct on (acts, labels, B, T, S, V), size and attach the workspace, optionally
    ``restrict_to_alignment``, then ``cost_and_grad``.  Reusable across calls on the same tensors.

    ``acts`` is either the reference's packed 2-D tensor or (extension) the joint network's own
    padded 4-D tensor [B, T, U, V] with U >= max S_b + 1: no gather into the packed layout, no scatter of the
    gradients back; gradients then have the same 4-D shape with exact zeros in the padding.
    """

    def __init__(self, acts: torch.Tensor, labels: torch.Tensor, input_lengths: torch.Tensor,
                 label_lengths: torch.Tensor, lengths_host: Optional[Tuple] = None):
        _check_inputs(acts, labels, input_lengths, label_lengths)
        self._lib = _lib.load()
        self.acts, self.labels = acts, labels
        self.input_lengths, self.label_lengths = input_lengths, label_lengths
        self.B = int(input_lengths.shape[0])
        self.V = int(acts.shape[-1])
        self.padded = acts.dim() == 4
        self._alignment = None
        if label_lengths.shape[0] != self.B or labels.shape[0] != self.B:
            raise ValueError("labels / input_lengths / label_lengths disagree on the batch size")
        th = sh = None
        if lengths_host is not None:
            self._T_h = np.ascontiguousarray(np.asarray(lengths_host[0], dtype=np.int32))
            self._S_h = np.ascontiguousarray(np.asarray(lengths_host[1], dtype=np.int32))
            th, sh = self._T_h.ctypes.data, self._S_h.ctypes.data
        h = ctypes.c_void_p()
        with torch.cuda.device(acts.device):
            if self.padded:
                st = self._lib.mrnnt_create_padded(ctypes.byref(h), acts.data_ptr(), labels.data_ptr(), self.B,
                                                   input_lengths.data_ptr(), label_lengths.data_ptr(), self.V,
                                                   int(acts.shape[1]), int(acts.shape[2]), int(labels.shape[1]),
                                                   th, sh)
                _lib.check(st, "mrnnt_create_padded")
            else:
                st = self._lib.mrnnt_create(ctypes.byref(h), acts.data_ptr(), labels.data_ptr(), self.B,
                                            input_lengths.data_ptr(), label_lengths.data_ptr(), self.V, th, sh)
                _lib.check(st, "mrnnt_create")
            self._h = h
            size = ctypes.c_size_t(0)
            if acts.dtype == torch.bfloat16:
                _lib.check(self._lib.mrnnt_set_dtype(self._h, 1), "mrnnt_set_dtype")
            _lib.check(self._lib.mrnnt_workspace_size(self._h, ctypes.byref(size)), "mrnnt_workspace_size")
            self.workspace_bytes = int(size.value)
            # The ABI derives the label stride (max_b S_b) and the packed row count from the LENGTHS, not from the
            # tensors (cpu_workspace_manager.h:44,117-135); the reference mis-indexes a wider labels tensor silently
            # (SURVEY appendix C-11).  Here the shapes are checked against what the engine has just derived.
            t_max, s_max, rows = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int64(0)
            _lib.check(self._lib.mrnnt_get_shape(self._h, ctypes.byref(t_max), ctypes.byref(s_max), ctypes.byref(rows)),
                       "mrnnt_get_shape")
            self.T_max, self.S_max, self.rows = int(t_max.value), int(s_max.value), int(rows.value)
            if not self.padded:
                if int(acts.shape[0]) != self.rows:
                    self.close()
                    raise ValueError(f"packed acts must have sum_b T_b*(S_b+1) = {self.rows} rows, got {int(acts.shape[0])}")
                if self.S_max >= 1 and (labels.dim() != 2 or int(labels.shape[1]) != self.S_max):
                    self.close()
                    raise ValueError(f"packed layout: labels must be [B, max_b S_b] = [{self.B}, {self.S_max}] (the ABI strides "
                                     f"it by max_b S_b), got {tuple(labels.shape)}")
            # caller-owned workspace, as the TensorFlow op does with allocate_temp (monotonic_rnnt_op.cu:117-123)
            self.workspace = torch.empty(self.workspace_bytes, dtype=torch.uint8, device=acts.device)
            _lib.check(self._lib.mrnnt_set_workspace(self._h, self.workspace.data_ptr()), "mrnnt_set_workspace")

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.mrnnt_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, option: int, value: int) -> None:
        _lib.check(self._lib.mrnnt_set_option(self._h, option, value), "mrnnt_set_option")

    def get_option(self, option: int) -> int:
        import ctypes
        v = ctypes.c_int(0)
        _lib.check(self._lib.mrnnt_get_option(self._h, option, ctypes.byref(v)), "mrnnt_get_option")
        return int(v.value)

    def set_peer_reduce(self, boards, total_out: Optional[torch.Tensor]) -> None:
        """From now on every call of this handle also leaves the sum over all ranks of the summed cost in
        ``total_out`` (one float32, on the device or in pinned host memory); see ``peer.PeerBoards`` and
        include/mrnnt_c_api.h (mrnnt_set_peer_reduce).  ``boards=None`` turns it off.  A handle that takes over boards
        another handle has used continues at ``boards.epoch``; ``sync_peer_epoch()`` writes it back."""
        if boards is None:
            _lib.check(self._lib.mrnnt_set_peer_reduce(self._h, 0, 0, None, None, 0), "mrnnt_set_peer_reduce")
            self._boards = None
            return
        if total_out is not None and (total_out.dtype != torch.float32 or total_out.numel() != 1):
            raise ValueError("total_out: one float32")
        self._boards, self._peer_total = boards, total_out
        _lib.check(self._lib.mrnnt_set_peer_reduce(self._h, boards.rank, boards.world, boards.c_array(),
                                                   None if total_out is None else total_out.data_ptr(), boards.epoch),
                   "mrnnt_set_peer_reduce")

    def set_peer_timeout_ms(self, ms: int) -> None:
        """How long a call waits for the slowest rank's cost sum (default 60 s, 0 = for ever); see mrnnt_c_api.h."""
        _lib.check(self._lib.mrnnt_set_peer_timeout_ms(self._h, int(ms)), "mrnnt_set_peer_timeout_ms")

    def peer_failed(self) -> bool:
        return bool(self._lib.mrnnt_peer_failed(self._h))

    def sync_peer_epoch(self) -> None:
        if getattr(self, "_boards", None) is not None:
            self._boards.epoch = int(self._lib.mrnnt_peer_epoch(self._h))

    def last_timings(self):
        """(ms_K1, ms_K2, ms_K3) of the last call; needs set_option(OPT_TIMING, 1) and a synchronised stream."""
        out = (ctypes.c_float * 3)()
        _lib.check(self._lib.mrnnt_last_timings(self._h, out), "mrnnt_last_timings")
        return float(out[0]), float(out[1]), float(out[2])

    def restrict_to_alignment(self, alignment: torch.Tensor, max_shift: int, blank_idx: int) -> None:
        if not alignment.is_cuda or alignment.dtype != torch.int32 or not alignment.is_contiguous():
            raise TypeError("alignment must be a contiguous CUDA int32 tensor [B, T_max]")
        if alignment.dim() != 2 or int(alignment.shape[0]) != self.B or int(alignment.shape[1]) < self.T_max:
            raise ValueError(f"alignment must be [B, >= max_b T_b] = [{self.B}, >= {self.T_max}], got {tuple(alignment.shape)}")
        self._alignment = alignment  # keep alive until the next compute call consumes it
        width = int(alignment.shape[1])
        if width == self.T_max:      # the reference's layout (cpu_workspace_manager.h:208)
            st = self._lib.mrnnt_restrict_to_alignment(self._h, alignment.data_ptr(), int(max_shift), int(blank_idx))
        else:                        # wider (e.g. [B, T_dim] next to a padded acts tensor): its own row stride
            st = self._lib.mrnnt_restrict_to_alignment_strided(self._h, alignment.data_ptr(), width, int(max_shift),
                                                               int(blank_idx))
        _lib.check(st, "mrnnt_restrict_to_alignment")

    def upload_acts(self, host_acts: torch.Tensor, stream: Optional[torch.cuda.Stream] = None) -> None:
        """Fill the device ``acts`` of this handle from a PINNED host tensor of the same shape and dtype, moving only
        the rows the lattice reads (mrnnt_upload_acts); asynchronous on ``stream``.  After restrict_to_alignment."""
        if host_acts.is_cuda or not host_acts.is_pinned() or not host_acts.is_contiguous():
            raise TypeError("host_acts must be a contiguous pinned host tensor")
        if host_acts.dtype != self.acts.dtype or host_acts.numel() != self.acts.numel():
            raise TypeError("host_acts must have the dtype and the size of acts")
        with self._device():
            st = stream if stream is not None else torch.cuda.current_stream(self.acts.device)
            _lib.check(self._lib.mrnnt_upload_acts(self._h, host_acts.data_ptr(), st.cuda_stream), "mrnnt_upload_acts")

    def _device(self):
        """Make acts' device current for the call; free when it already is (the usual case)."""
        idx = self.acts.device.index
        if idx is None or torch.cuda.current_device() == idx:
            return _NULL_CONTEXT
        return torch.cuda.device(idx)

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.acts.device).cuda_stream

    def cost_and_grad(self, blank_label: int = 0, grads: Optional[torch.Tensor] = None,
                      costs_host: Optional[torch.Tensor] = None) -> torch.Tensor:
        """GpuRNNTComputer::cost_and_grad: costs land on the HOST, valid on return.  With ``grads`` the call returns as
        soon as the costs have arrived; the gradient kernel may still be running and ``grads`` is complete in stream
        order, like the result of any kernel launch (``set_option(OPT_RETURN_EARLY, 0)``: wait for everything)."""
        if costs_host is None:
            costs_host = torch.empty(self.B, dtype=torch.float32, device="cpu")
        assert costs_host.device.type == "cpu" and costs_host.dtype == torch.float32 and costs_host.numel() == self.B
        gptr = None
        if grads is not None:
            assert grads.is_cuda and grads.dtype == self.acts.dtype and grads.is_contiguous()
            assert grads.numel() == self.acts.numel()
            gptr = grads.data_ptr()
        with self._device():
            st = self._lib.mrnnt_cost_and_grad(self._h, int(blank_label), self._stream(), costs_host.data_ptr(), gptr)
        _lib.check(st, "mrnnt_cost_and_grad")
        return costs_host

    def cost(self, blank_label: int = 0, costs_host: Optional[torch.Tensor] = None) -> torch.Tensor:
        return self.cost_and_grad(blank_label, None, costs_host)

    def enqueue(self, blank_label: int = 0, grads: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Launch without synchronising the host; returns a device VIEW of the costs (valid in stream order)."""
        gptr = grads.data_ptr() if grads is not None else None
        with self._device():
            st = self._lib.mrnnt_enqueue(self._h, int(blank_label), self._stream(), gptr)
        _lib.check(st, "mrnnt_enqueue")
        return self.device_costs()

    def enqueue_forward(self, blank_label: int = 0, want_grads: bool = True,
                        grads: Optional[torch.Tensor] = None) -> torch.Tensor:
        """First half of a call (K1 + K2); returns a device VIEW of the costs.  With want_grads the workspace
        keeps what enqueue_backward() needs; acts and the workspace must stay untouched until then.  `grads`
        (optional): the buffer enqueue_backward() will be given -- its zero rows are then written here, while the
        lattice recursions run; it must not be written in between."""
        with self._device():
            if grads is not None:
                assert grads.is_cuda and grads.dtype == self.acts.dtype and grads.is_contiguous()
                assert grads.numel() == self.acts.numel()
                st = self._lib.mrnnt_enqueue_forward_into(self._h, int(blank_label), self._stream(), grads.data_ptr())
            else:
                st = self._lib.mrnnt_enqueue_forward(self._h, int(blank_label), self._stream(), 1 if want_grads else 0)
        _lib.check(st, "mrnnt_enqueue_forward")
        return self.device_costs()

    def enqueue_backward(self, grads: torch.Tensor, scale: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Second half (K3): grads[i, :] = scale[b(i)] * d cost_b / d acts[i, :], written exactly once."""
        assert grads.is_cuda and grads.dtype == self.acts.dtype and grads.is_contiguous()
        assert grads.numel() == self.acts.numel()
        sptr = None
        if scale is not None:
            assert scale.is_cuda and scale.dtype == torch.float32 and scale.is_contiguous() and scale.numel() == self.B
            sptr = scale.data_ptr()
        with self._device():
            st = self._lib.mrnnt_enqueue_backward(self._h, self._stream(), grads.data_ptr(), sptr)
        _lib.check(st, "mrnnt_enqueue_backward")
        return grads

    def device_costs(self) -> torch.Tensor:
        ptr = self._lib.mrnnt_device_costs(self._h)
        off = int(ptr) - self.workspace.data_ptr()
        assert 0 <= off <= self.workspace_bytes - 4 * self.B
        return self.workspace[off:off + 4 * self.B].view(torch.float32)

    def debug(self, what: int) -> np.ndarray:
        """Intermediate arrays for kernel-level parity tests (blocking)."""
        T = self.input_lengths.cpu().numpy().astype(np.int64)
        S = self.label_lengths.cpu().numpy().astype(np.int64)
        rows = int(self.acts.numel() // self.V) if self.padded else int((T * (S + 1)).sum())
        shapes = {
            _lib.DBG_DENOM: (np.float64, (rows,)), _lib.DBG_ALPHA: (np.float64, (rows,)),
            _lib.DBG_BETA: (np.float64, (rows,)), _lib.DBG_LP: (np.float64, (rows, 2)),
            _lib.DBG_BAND: (np.int32, (self.B, int(T.max()), 2)), _lib.DBG_ROWMETA: (np.int32, (rows,)),
            _lib.DBG_LL: (np.float64, (2, self.B)), _lib.DBG_ROWSTART: (np.int64, (self.B + 1,)),
        }
        dt, shape = shapes[what]
        out = np.empty(shape, dtype=dt)
        with self._device():
            _lib.check(self._lib.mrnnt_debug_copy(self._h, what, out.ctypes.data, out.nbytes), "mrnnt_debug_copy")
        return out


class MonotonicRNNTFunction(torch.autograd.Function):
    """Same call signature as the reference's autograd function (monotonic_rnnt_op.py:19-118).

    The reference computes the gradients inside forward (into a zeros_like buffer) and multiplies them by the
    upstream gradient in backward: a memset, a read and a write of logits-sized arrays on top of the loss
    itself.  Here forward runs K1 + K2 only and backward runs K3 with the upstream gradient folded into its
    single write, so a training step moves 12 bytes per logit in total and an inference call (no backward)
    never touches the gradient array at all.
    """

    @staticmethod
    def forward(ctx, acts, labels, input_lengths, label_lengths, alignment=None, max_distance_from_alignment=0,
                blank_label=0):
        handle = LossHandle(acts, labels, input_lengths, label_lengths)
        if alignment is not None:
            handle.restrict_to_alignment(alignment, max_distance_from_alignment, blank_label)
        want_grads = bool(acts.requires_grad)
        costs = handle.enqueue_forward(blank_label, want_grads).clone()
        stream = torch.cuda.current_stream(acts.device)
        # the workspace and alignment must outlive the kernels that are still in flight on this stream
        handle.workspace.record_stream(stream)
        if want_grads:
            ctx.handle = handle          # keeps the workspace (coefficients) alive until backward
            ctx.save_for_backward(acts)  # backward re-streams the logits: they must not be modified in between
        else:
            handle.close()
        return costs

    @staticmethod
    def backward(ctx, grad_outputs):
        # every logit of utterance b receives d cost_b / d logit * grad_outputs[b] (monotonic_rnnt_op.py:97-118)
        (acts,) = ctx.saved_tensors
        handle = ctx.handle
        grads = torch.empty_like(acts)
        scale = grad_outputs.detach().to(device=acts.device, dtype=torch.float32).contiguous()
        handle.enqueue_backward(grads, scale)
        stream = torch.cuda.current_stream(acts.device)
        handle.workspace.record_stream(stream)
        scale.record_stream(stream)
        return grads, None, None, None, None, None, None


def monotonic_rnnt_loss(acts, labels, input_lengths, label_lengths, alignment: Optional[torch.Tensor] = None,
                        max_distance_from_alignment: int = 0, blank_label: int = 0) -> torch.Tensor:
    """Monotonic RNN-T loss per utterance (negative log-likelihood), softmax applied internally.

    Arguments as in the reference (monotonic_rnnt_op.py:121-152): ``acts`` packed
    [sum_b T_b*(S_b+1), V] float32, ``labels`` [B, max_b S_b] int32, ``input_lengths`` / ``label_lengths``
    [B] int32, optional ``alignment`` [B, max_b T_b] int32 with ``max_distance_from_alignment``.
    Extension: ``acts`` may be the padded joint-network output [B, T, U, V] (then ``labels`` is [B, S] with any
    S >= max_b S_b); its gradient has the same shape, zero in the padding.
    Returns a float32 tensor [B] on ``acts.device``.
    """
    result = MonotonicRNNTFunction.apply(acts, labels, input_lengths, label_lengths, alignment,
                                         max_distance_from_alignment, blank_label)
    assert result is not None
    return result


class MonotonicRNNTLoss(torch.nn.Module):
    """Module form (reference monotonic_rnnt_op.py:166-217)."""

    def __init__(self, blank_label: int = 0) -> None:
        super().__init__()
        self.blank_label = blank_label

    def forward(self, acts, labels, input_lengths, label_lengths, alignment: Optional[torch.Tensor] = None,
                max_distance_from_alignment: int = 0) -> torch.Tensor:
        return monotonic_rnnt_loss(acts, labels, input_lengths, label_lengths, alignment,
                                   max_distance_from_alignment, self.blank_label)
