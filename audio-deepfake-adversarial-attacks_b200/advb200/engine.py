"""Host-side engine handles: one ``libadvb200`` handle per (model instance, device, clip length).

Everything numerical happens behind the C ABI; torch is used for device memory, the current CUDA stream and the
random-start noise only (parity needs torch's RNG).
"""
import ctypes as C
import weakref

import torch
from torch import nn

from . import _lib

_MODEL_KINDS = {"LCNN": _lib.MODEL_LCNN, "SpecRNet": _lib.MODEL_SPECRNET, "RawNet3": _lib.MODEL_RAWNET3}
_ENGINES = weakref.WeakKeyDictionary()  # module -> {(device index, T): Engine}


def unwrap(model: nn.Module) -> nn.Module:
    """The reference hands the attack an ``nn.DataParallel`` (evaluate_models_on_adversarial_attacks.py:163-169)."""
    return model.module if isinstance(model, nn.DataParallel) else model


def _frontend_kind(state) -> int:
    if "frontend.filter_mat" in state:
        return _lib.FRONTEND_LFCC
    if "frontend.MelSpectrogram.mel_scale.fb" in state:
        return _lib.FRONTEND_MFCC
    return _lib.FRONTEND_NONE


def _require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise RuntimeError(f"advb200 has no CPU path: {what} is on {t.device}; move the model and batch to a CUDA device")


def _is_frontend(module: nn.Module) -> bool:
    """A bare LFCC / MFCC transform (torchaudio's or the advb200 holder): ``LFCC_FN`` / ``MFCC_FN`` of src/frontends.py."""
    keys = dict(module.named_buffers()).keys()
    return "dct_mat" in keys and ("filter_mat" in keys or "MelSpectrogram.mel_scale.fb" in keys)


class Engine:
    def __init__(self, module: nn.Module, max_batch: int, n_samples: int):
        self.lib = _lib.load()
        self._prefix = ""
        kind = _MODEL_KINDS.get(type(module).__name__)
        if kind is None and _is_frontend(module):
            kind, self._prefix = _lib.MODEL_FRONTEND_ONLY, "frontend."
        if kind is None:
            raise NotImplementedError(f"advb200 engine does not support model class {type(module).__name__}")
        self.module_ref = weakref.ref(module)
        self.max_batch, self.n_samples = int(max_batch), int(n_samples)
        refs, self._ptrs = self._tensor_table(module)
        first = next(iter(self._state.values()))
        _require_cuda(first, "the model")
        self.device = first.device
        desc = _lib.ModelDesc(kind, _frontend_kind(self._state), self.device.index or 0, self.max_batch, self.n_samples,
                              len(refs), refs)
        handle = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.advb_create(C.byref(handle), C.byref(desc)))
        self.handle = handle
        self._finalizer = weakref.finalize(self, self.lib.advb_destroy, handle)
        # this host tracks weight versions itself (data_ptr + tensor._version stamps): unchanged weights skip the repack
        _lib.check(self.lib.advb_set_option(handle, b"weight_cache", 1))
        self._versions = None

    def _tensor_table(self, module):
        live = {self._prefix + k: v for k, v in module.state_dict(keep_vars=True).items() if v.dtype == torch.float32}
        state, self._mirrors = {}, {}
        for k, v in live.items():
            _require_cuda(v, f"tensor '{k}'")
            if not v.is_contiguous():
                # torchaudio registers LFCC / MFCC's dct_mat as a transposed VIEW (create_dct(...) returns dct.t()): the
                # engine reads a packed copy, refreshed whenever the live tensor's version counter moves (_sync_weights)
                self._mirrors[k] = (v, v.detach().contiguous())
                state[k] = self._mirrors[k][1]
            else:
                state[k] = v
        self._state = state  # keeps the storages alive while the handle borrows them
        self._names = [k.encode() for k in state]
        refs = (_lib.TensorRef * len(state))()
        for i, (k, v) in enumerate(state.items()):
            refs[i] = _lib.TensorRef(self._names[i], v.data_ptr(), v.numel())
        return refs, tuple(v.data_ptr() for v in live.values())  # the LIVE pointers: what _sync_weights compares

    def _sync_weights(self):
        """Weights are read live.  A storage that was *replaced* is re-pointed (load_state_dict keeps storages); a tensor
        whose autograd version counter moved (optimizer.step(), load_state_dict, any in-place op on the parameter) marks the
        packed weight images stale.  In-place writes that bypass the version counter (``p.data.mul_()``) are invisible to
        this check: call ``engine.invalidate()`` after them."""
        module = self.module_ref()
        if module is None:
            raise RuntimeError("model was garbage-collected")
        live = [v for v in module.state_dict(keep_vars=True).values() if v.dtype == torch.float32]
        ptrs = tuple(v.data_ptr() for v in live)
        if ptrs != self._ptrs:
            refs, self._ptrs = self._tensor_table(module)
            _lib.check(self.lib.advb_rebind(self.handle, len(refs), refs))
            self._versions = None
        versions = tuple(v._version for v in live)
        if versions != self._versions:
            for src, packed in self._mirrors.values():
                packed.copy_(src.detach())
            _lib.check(self.lib.advb_invalidate_weights(self.handle))
            self._versions = versions

    def invalidate(self):
        """Force the next call to repack the weights (after writes the version stamps cannot see)."""
        self._versions = None

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _check_x(self, x):
        _require_cuda(x, "the batch")
        if x.dtype != torch.float32 or x.dim() != 2:
            raise ValueError("expected a float32 waveform batch of shape (B, T)")
        if x.shape[1] != self.n_samples or x.shape[0] > self.max_batch:
            raise ValueError("batch does not fit this engine handle")
        return x.contiguous()

    def attack(self, desc: "_lib.AttackDesc", x, y, start=None, minmax=False, target=None):
        """``minmax=True``: ``x`` is the raw waveform batch; to_minmax / revert_minmax run inside the same native call.
        ``target``: target labels of the targeted modes (attack.py:60-108); sets ``desc.targeted``."""
        x = self._check_x(x)
        y = self._check_labels(y, x)
        self._sync_weights()
        out = torch.empty_like(x)
        sp = None
        if start is not None:
            if start.dtype != torch.float32 or start.shape != x.shape or start.device != x.device:
                raise ValueError(f"random start must be a float32 tensor of shape {tuple(x.shape)} on {x.device}, got "
                                 f"{start.dtype} {tuple(start.shape)} on {start.device}")
            start = start.contiguous()  # kept alive in this frame until the call has been enqueued
            sp = C.c_void_p(start.data_ptr())
        desc.targeted, desc.target_labels = 0, None
        if target is not None:
            target = self._check_labels(target, x)  # kept alive in this frame until the call has been enqueued
            desc.targeted, desc.target_labels = 1, target.data_ptr()
        fn = self.lib.advb_attack_minmax if minmax else self.lib.advb_attack
        with torch.cuda.device(self.device):
            _lib.check(fn(self.handle, C.byref(desc), x.data_ptr(), y.data_ptr(), sp, out.data_ptr(),
                                            x.shape[0], x.shape[1], self._stream()))
        return out

    @staticmethod
    def _check_labels(y, x):
        """int64 labels in {0, 1} on the batch's device (the head uses (float) y and FAB / CW assume two classes)."""
        if y is None or y.dim() != 1 or y.shape[0] != x.shape[0]:
            raise ValueError(f"expected {x.shape[0]} labels of shape (B,)")
        if y.dtype.is_floating_point or y.dtype == torch.bool:
            raise ValueError(f"labels must be integer class indices, got {y.dtype}")
        if not y.is_cuda:  # free on the host (src/trainer.py passes CPU labels)
            if y.numel() and (int(y.min()) < 0 or int(y.max()) > 1):
                raise ValueError("labels must be 0 (spoof) or 1 (bonafide)")
        else:  # no host sync: device-side assertion
            torch._assert_async(((y == 0) | (y == 1)).all())
        return y.to(device=x.device, dtype=torch.int64).contiguous()

    def forward(self, x):
        x = self._check_x(x)
        self._sync_weights()
        out = torch.empty(x.shape[0], 1, device=x.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.advb_forward(self.handle, x.data_ptr(), out.data_ptr(), x.shape[0], x.shape[1],
                                             self._stream()))
        return out

    def grad(self, x, y=None, what=_lib.GRAD_CE, n_global=0):
        x = self._check_x(x)
        self._sync_weights()
        g = torch.empty_like(x)
        logits = torch.empty(x.shape[0], 1, device=x.device, dtype=torch.float32)
        yp = None
        if y is not None:
            y = y.to(device=x.device, dtype=torch.int64).contiguous()
            yp = C.c_void_p(y.data_ptr())
        with torch.cuda.device(self.device):
            _lib.check(self.lib.advb_grad(self.handle, what, x.data_ptr(), yp, g.data_ptr(), logits.data_ptr(),
                                          x.shape[0], x.shape[1], n_global, self._stream()))
        return g, logits

    def frontend_fwd(self, x):
        x = self._check_x(x)
        self._sync_weights()
        F = 1 + x.shape[1] // 160
        out = torch.empty(x.shape[0], 80, F, device=x.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.advb_frontend_fwd(self.handle, x.data_ptr(), out.data_ptr(), x.shape[0], x.shape[1],
                                                  self._stream()))
        return out

    def frontend_bwd(self, x, g_coeff):
        x = self._check_x(x)
        g_coeff = g_coeff.contiguous()
        gx = torch.empty_like(x)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.advb_frontend_bwd(self.handle, x.data_ptr(), g_coeff.data_ptr(), gx.data_ptr(),
                                                  x.shape[0], x.shape[1], self._stream()))
        return gx

    def row_diff_norms(self, a, b):
        """Per-clip (L-inf, L2) norms of ``a - b`` (fab.py:515-521)."""
        a, b = a.contiguous(), b.contiguous()
        linf = torch.empty(a.shape[0], device=a.device, dtype=torch.float32)
        l2 = torch.empty_like(linf)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.advb_row_diff_norms(a.data_ptr(), b.data_ptr(), linf.data_ptr(), l2.data_ptr(),
                                                    a.shape[0], a.shape[1], self._stream()))
        return linf, l2

    def debug_stage(self, name: str):
        """(tensor (B, H+2p, W+2p, C) raw engine layout, pad p) of an internal stage of the last forward."""
        dims = (C.c_int64 * 5)()
        n = self.lib.advb_debug_stage(self.handle, name.encode(), None, 0, dims, self._stream())
        if n < 0:
            _lib.check(1)
        buf = torch.empty(n, device=self.device, dtype=torch.float32)
        n = self.lib.advb_debug_stage(self.handle, name.encode(), buf.data_ptr(), n, dims, self._stream())
        if n < 0:
            _lib.check(1)
        B, H, W, Cc, p = [int(v) for v in dims]
        return buf.view(B, H + 2 * p, W + 2 * p, Cc), p

    def set_option(self, key: str, value: int):
        """``conv_path`` (0 tcgen05 / 1 fp32 SIMT), ``tf32_passes`` (3 = 3xTF32 / 1 = single pass), ``conv_sched``
        (0 persistent kernels / 1 one-tile-per-CTA kernels), ``conv0_bwd`` (first block backward: 0 fp32 cell kernel /
        1 tcgen05 GEMM + col2im), ``graph`` (1 = replay the attack iteration as a CUDA graph), ``fuse_update`` (1 = FGSM /
        PGD update rule in the frontend backward's epilogue), ``sr_tc`` (SpecRNet 3x3 convolutions: 1 tcgen05 / 0 fp32 SIMT),
        ``fe_spec``, ``lstm_tc``, ``conv0_fwd``, ``weight_cache``: see ``include/advb200.h``."""
        _lib.check(self.lib.advb_set_option(self.handle, key.encode(), int(value)))

    # ---- strict multi-GPU mode (include/advb200.h, advb_xrank_*; SURVEY.md §8(e) ii) ---------------------------------
    def xrank_export(self):
        """(64-byte CUDA IPC handle, device pointer) of this handle's mailbox; clears the mailbox (protocol step 1)."""
        buf = C.create_string_buffer(_lib.XRANK_HANDLE_BYTES)
        ptr = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.advb_xrank_export(self.handle, buf, C.byref(ptr)))
        return buf.raw, ptr.value

    def xrank_connect(self, rank: int, world: int, ipc_handles=None, local_ptrs=None):
        """Map the peers' mailboxes (protocol step 3).  ``ipc_handles``: the ranks' 64-byte handles in rank order (peers in
        other processes); ``local_ptrs``: device pointers of handles living in this process (``None`` entries fall back to
        the IPC handle).  ``world <= 1`` returns to per-shard floors."""
        blob = b"".join(ipc_handles) if ipc_handles is not None else None
        if blob is not None and len(blob) != world * _lib.XRANK_HANDLE_BYTES:
            raise ValueError("one 64-byte IPC handle per rank, in rank order")
        ptrs = None
        if local_ptrs is not None:
            ptrs = (C.c_void_p * world)(*[C.c_void_p(p) if p else C.c_void_p() for p in local_ptrs])
        with torch.cuda.device(self.device):
            _lib.check(self.lib.advb_xrank_connect(self.handle, int(rank), int(world), blob, ptrs))
        self.strict_world = int(world) if world > 1 else 1

    def enable_strict(self, group=None):
        """One process per GPU: from now on the dB floor of every frontend pass spans the clips of ALL ranks of ``group``, so
        the sharded run reproduces the single-device batch.  Collective: every rank of the group must call it, and must then
        make the same sequence of forward / gradient / FGSM / PGD / PGDL2 calls."""
        import torch.distributed as dist

        world, rank = dist.get_world_size(group), dist.get_rank(group)
        if world == 1:
            return self.xrank_connect(0, 1)
        handle, _ = self.xrank_export()
        handles = [None] * world
        dist.all_gather_object(handles, handle, group=group)  # also the barrier between "cleared" and "first exchange"
        self.xrank_connect(rank, world, ipc_handles=handles)
        dist.barrier(group=group)  # nobody starts exchanging before every rank has mapped its peers

    def strict_timed_out(self) -> bool:
        """True when an exchange gave up waiting for a peer (the ranks' call sequences diverged, or a peer died); the results
        since then used per-shard floors.  Synchronises the device."""
        torch.cuda.synchronize(self.device)
        flag = C.c_int(0)
        _lib.check(self.lib.advb_xrank_status(self.handle, C.byref(flag)))
        return bool(flag.value)

    def profile_begin(self):
        _lib.check(self.lib.advb_profile_begin(self.handle, self._stream()))

    def profile_end(self):
        """[{name, count, total_ms}] per kernel since profile_begin (synchronises)."""
        import json

        buf = C.create_string_buffer(1 << 16)
        n = self.lib.advb_profile_end(self.handle, buf, len(buf))
        if n < 0 or n > len(buf):
            raise RuntimeError("advb_profile_end failed")
        return json.loads(buf.value.decode() or "[]")

    @property
    def launches(self) -> int:
        return int(self.lib.advb_launch_count(self.handle))

    @property
    def workspace_bytes(self) -> int:
        return int(self.lib.advb_workspace_bytes(self.handle))


def engine_for(model: nn.Module, batch: int, n_samples: int) -> Engine:
    module = unwrap(model)
    first = next(module.parameters(), None)
    if first is None:  # a bare frontend transform has buffers only
        first = next(module.buffers())
    _require_cuda(first, "the model")
    per_module = _ENGINES.setdefault(module, {})
    key = (first.device.index or 0, int(n_samples))
    eng = per_module.get(key)
    if eng is None or eng.max_batch < batch:
        if eng is not None:
            eng._finalizer()
        eng = Engine(module, max(batch, eng.max_batch if eng else 0), n_samples)
        per_module[key] = eng
    return eng


def model_forward(module: nn.Module, x: torch.Tensor) -> torch.Tensor:
    """``model(x)`` of an advb200 model holder: logits (B,1), computed by the CUDA engine."""
    _require_cuda(x, "the batch")
    return engine_for(module, x.shape[0], x.shape[1]).forward(x)


def frontend_forward(frontend: nn.Module, x: torch.Tensor) -> torch.Tensor:
    """``LFCC_FN(x)`` / ``MFCC_FN(x)`` (src/frontends.py:13-32): coefficients (B, 80, 1 + T // 160), computed by the CUDA
    frontend kernels on a frontend-only engine handle that borrows the transform's live buffers."""
    _require_cuda(x, "the batch")
    squeeze = x.dim() == 1
    if squeeze:
        x = x.unsqueeze(0)
    out = engine_for(frontend, x.shape[0], x.shape[1]).frontend_fwd(x)
    return out[0] if squeeze else out


def projection_linf(t: torch.Tensor, w: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """fab.py:562-614 on the GPU (rows independent)."""
    _require_cuda(t, "the points")
    lib = _lib.load()
    t, w, b = t.contiguous(), w.contiguous(), b.contiguous()
    d = torch.empty_like(t)
    with torch.cuda.device(t.device):
        _lib.check(lib.advb_projection_linf(t.data_ptr(), w.data_ptr(), b.data_ptr(), d.data_ptr(), t.shape[0], t.shape[1],
                                            C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)))
    return d


def projection_l2(t: torch.Tensor, w: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """fab.py:617-665 on the GPU (rows independent)."""
    _require_cuda(t, "the points")
    lib = _lib.load()
    t, w, b = t.contiguous(), w.contiguous(), b.contiguous()
    d = torch.empty_like(t)
    with torch.cuda.device(t.device):
        _lib.check(lib.advb_projection_l2(t.data_ptr(), w.data_ptr(), b.data_ptr(), d.data_ptr(), t.shape[0], t.shape[1],
                                          C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)))
    return d


def mel_spec_forward(audio: torch.Tensor, fb: torch.Tensor) -> torch.Tensor:
    """src/frontends.py:53-79 on the GPU: (B, 2, n_mels, F) = [abs, angle] of the mel-scaled complex STFT.  Forward only."""
    _require_cuda(audio, "the waveform")
    if audio.requires_grad:
        raise NotImplementedError("the mel_spec frontend is forward-only in advb200 (no model of the reference consumes it)")
    squeeze = audio.dim() == 1
    x = (audio.unsqueeze(0) if squeeze else audio).contiguous().float()
    fbc = fb.to(x.device).contiguous().float()
    if fbc.dim() != 2 or fbc.shape[0] != 257:
        raise ValueError("mel filterbank must be (257, n_mels)")
    lib = _lib.load()
    B, T = x.shape
    out = torch.empty(B, 2, fbc.shape[1], 1 + T // 160, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        _lib.check(lib.advb_mel_spec_fwd(x.data_ptr(), fbc.data_ptr(), fbc.shape[1], out.data_ptr(), B, T,
                                         C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)))
    return out[0] if squeeze else out


def to_minmax(x: torch.Tensor):
    """src/aa/utils.py:4-9 on the GPU: returns (x01, mn (B,1), mx (B,1))."""
    _require_cuda(x, "the batch")
    lib = _lib.load()
    x = x.contiguous()
    out = torch.empty_like(x)
    mn = torch.empty(x.shape[0], 1, device=x.device, dtype=torch.float32)
    mx = torch.empty_like(mn)
    with torch.cuda.device(x.device):
        _lib.check(lib.advb_minmax(x.data_ptr(), out.data_ptr(), mn.data_ptr(), mx.data_ptr(), x.shape[0], x.shape[1],
                                   C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)))
    return out, mn, mx


def revert_minmax(x01: torch.Tensor, mn: torch.Tensor, mx: torch.Tensor):
    """src/aa/utils.py:12-14 on the GPU."""
    _require_cuda(x01, "the batch")
    lib = _lib.load()
    x01 = x01.contiguous()
    out = torch.empty_like(x01)
    with torch.cuda.device(x01.device):
        _lib.check(lib.advb_revert_minmax(x01.data_ptr(), mn.contiguous().data_ptr(), mx.contiguous().data_ptr(),
                                          out.data_ptr(), x01.shape[0], x01.shape[1],
                                          C.c_void_p(torch.cuda.current_stream(x01.device).cuda_stream)))
    return out
