"""ctypes binding of libsr3_b200.so (include/sr3_b200.h).  torch is used only for device memory and streams."""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int64, c_uint64, c_void_p

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "lib", "libsr3_b200.so")
SR3_MAX_LEVELS = 8
_lib = None


class NativeLibraryError(RuntimeError):
    pass


class UNetConfigC(ctypes.Structure):
    _fields_ = [("in_channel", c_int), ("out_channel", c_int), ("inner_channel", c_int), ("norm_groups", c_int),
                ("n_mults", c_int), ("channel_mults", c_int * SR3_MAX_LEVELS), ("n_attn_res", c_int),
                ("attn_res", c_int * SR3_MAX_LEVELS), ("res_blocks", c_int), ("image_size", c_int), ("channels", c_int),
                ("conditional", c_int), ("precision", c_int)]


PRECISIONS = {"bf16": 0, "fp32": 1}          # "fp32" = precise mode: (hi, lo) bf16 operand pairs, three tensor-core passes


_SIGS = {
    "sr3_last_error": (c_char_p, []),
    "sr3_abi_version": (c_int, []),
    "sr3_engine_create": (c_int, [POINTER(UNetConfigC), c_int, c_int, POINTER(c_void_p)]),
    "sr3_engine_create_train": (c_int, [POINTER(UNetConfigC), c_int, c_int, c_float, POINTER(c_void_p)]),
    "sr3_train_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_uint64, POINTER(c_double), c_void_p]),
    "sr3_train_backward": (c_int, [c_void_p, c_float, POINTER(c_void_p), c_int, c_void_p]),
    "sr3_train_num_backward_blocks": (c_int, [c_void_p]),
    "sr3_train_backward_begin": (c_int, [c_void_p, c_float, POINTER(c_void_p), c_int]),
    "sr3_train_backward_block": (c_int, [c_void_p, c_int, c_void_p]),
    "sr3_train_backward_flush": (c_int, [c_void_p, c_void_p]),
    "sr3_train_backward_finish": (c_int, [c_void_p, c_void_p]),
    "sr3_train_block_params": (c_int, [c_void_p, c_int, POINTER(c_int), c_int, POINTER(c_int)]),
    "sr3_train_backward_profile": (c_int, [c_void_p, c_float, POINTER(c_void_p), c_int, POINTER(c_float), c_void_p]),
    "sr3_train_set_dropout_mask": (c_int, [c_void_p, c_char_p, c_void_p]),
    "sr3_train_num_dropout_layers": (c_int, [c_void_p]),
    "sr3_train_dropout_layer_name": (c_int, [c_void_p, c_int, c_char_p, c_int]),
    "sr3_adam_step": (c_int, [c_void_p, c_int, c_float, c_float, c_float, c_float, c_int, c_float, c_void_p]),
    "sr3_engine_destroy": (None, [c_void_p]),
    "sr3_engine_num_params": (c_int, [c_void_p]),
    "sr3_engine_param_info": (c_int, [c_void_p, c_int, c_char_p, c_int, POINTER(c_int64), POINTER(c_int)]),
    "sr3_engine_load_param": (c_int, [c_void_p, c_char_p, c_void_p, c_int64, c_void_p]),
    "sr3_engine_load_all_params": (c_int, [c_void_p, POINTER(c_void_p), c_int, c_void_p]),
    "sr3_engine_finalize_params": (c_int, [c_void_p, c_void_p]),
    "sr3_engine_set_schedule": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "sr3_unet_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "sr3_p_mean_variance": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, POINTER(c_float), c_void_p]),
    "sr3_p_sample": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_uint64, c_uint64, c_void_p, c_void_p]),
    "sr3_p_losses": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, POINTER(c_double), c_void_p]),
    "sr3_p_sample_loop": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_uint64, c_uint64, c_void_p, c_void_p, c_int,
                                  POINTER(c_int), c_void_p]),
    "sr3_super_resolution_host": (c_int, [c_void_p, c_void_p, c_void_p, c_uint64, c_uint64, c_void_p, c_void_p]),
    "sr3_p_sample_loop_begin": (c_int, [c_void_p, c_void_p, c_void_p, c_uint64, c_uint64, c_void_p]),
    "sr3_p_sample_steps": (c_int, [c_void_p, c_int, c_int, c_void_p]),
    "sr3_read_state": (c_int, [c_void_p, c_void_p, c_void_p]),
    "sr3_engine_profile_step": (c_int, [c_void_p, c_int, c_int, c_int, POINTER(c_int), POINTER(c_float), POINTER(c_double), POINTER(c_double),
                                        POINTER(c_int), c_void_p]),
    "sr3_pil_bicubic_tables": (c_int, [c_int, c_int, POINTER(c_int), POINTER(c_int), c_int, POINTER(c_int)]),
    "sr3_resize_bicubic_u8": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_float, c_void_p]),
    "sr3_tensor2img": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_float, c_void_p]),
    "sr3_ssd_u8": (c_int, [c_void_p, c_void_p, c_int64, POINTER(c_uint64), c_void_p]),
    "sr3_engine_num_launches_per_step": (c_int, [c_void_p]),
    "sr3_engine_num_ops_per_step": (c_int, [c_void_p]),
    "sr3_engine_uses_step_kernel": (c_int, [c_void_p]),
    "sr3_engine_step_kernel_profile": (c_int, [c_void_p, c_int, POINTER(c_int), POINTER(c_double), POINTER(c_double), POINTER(c_int), c_void_p]),
    "sr3_engine_workspace_bytes": (c_int64, [c_void_p]),
    "sr3_engine_read_activation": (c_int, [c_void_p, c_char_p, c_void_p, c_int64, POINTER(c_int64), POINTER(c_int), c_void_p]),
    "sr3_bench_conv": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, POINTER(c_float)]),
    "sr3_test_gemm": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "sr3_test_attention": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "sr3_test_conv_groupnorm": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                        c_int, c_void_p]),
    "sr3_test_conv": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                              c_void_p]),
}
EXPORTED_SYMBOLS = tuple(_SIGS.keys())


def lib():
    """Load (never build silently on a GPU box) the native library; raise loudly if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryError(
            f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` (nvcc, sm_100a). "
            "sr3_b200 has no CPU or eager-PyTorch fallback.")
    try:
        l = ctypes.CDLL(LIB_PATH)
    except OSError as e:
        raise NativeLibraryError(f"cannot load {LIB_PATH}: {e}") from e
    for name, (res, args) in _SIGS.items():
        try:
            fn = getattr(l, name)
        except AttributeError as e:
            raise NativeLibraryError(f"{LIB_PATH} does not export {name}") from e
        fn.restype = res
        fn.argtypes = args
    _lib = l
    return l


def _check(rc):
    if rc != 0:
        raise RuntimeError("sr3_b200: " + lib().sr3_last_error().decode())


def _stream():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return c_void_p(0) if t is None else c_void_p(t.data_ptr())


def _f32c(t, device):
    return t.detach().to(device=device, dtype=torch.float32).contiguous()


class Engine:
    """One (config, batch, device) instance of the native plan: packed weights + activations + captured step graph."""

    def __init__(self, cfg: dict, batch: int, device: torch.device, train_dropout=None):
        """train_dropout: None = inference plan; a float = TRAINING plan (forward keeps every intermediate, backward recorded) with that
        Dropout probability (sr3_engine_create_train)."""
        if device.type != "cuda":
            raise NativeLibraryError("sr3_b200 runs on a CUDA (sm_100a) device only; got device=%s" % device)
        self.device = device
        self.batch = batch
        self.channels = cfg["channels"]
        self.in_channel = cfg["in_channel"]
        self.out_channel = cfg["out_channel"]
        self.image_size = cfg["image_size"]
        self.conditional = bool(cfg["conditional"])
        c = UNetConfigC()
        c.in_channel, c.out_channel, c.inner_channel = cfg["in_channel"], cfg["out_channel"], cfg["inner_channel"]
        c.norm_groups = cfg["norm_groups"]
        mults, attn = list(cfg["channel_mults"]), list(cfg["attn_res"])
        c.n_mults, c.n_attn_res = len(mults), len(attn)
        for i, m in enumerate(mults):
            c.channel_mults[i] = m
        for i, a in enumerate(attn):
            c.attn_res[i] = a
        c.res_blocks, c.image_size, c.channels, c.conditional = cfg["res_blocks"], cfg["image_size"], cfg["channels"], int(cfg["conditional"])
        self.precision = cfg.get("precision", "bf16")
        if self.precision not in PRECISIONS:
            raise ValueError("precision must be one of %s, got %r" % (sorted(PRECISIONS), self.precision))
        c.precision = PRECISIONS[self.precision]
        self._h = c_void_p()
        idx = device.index if device.index is not None else torch.cuda.current_device()
        self.train_dropout = train_dropout
        if train_dropout is None:
            _check(lib().sr3_engine_create(ctypes.byref(c), batch, idx, ctypes.byref(self._h)))
        else:
            _check(lib().sr3_engine_create_train(ctypes.byref(c), batch, idx, float(train_dropout), ctypes.byref(self._h)))
        self.T = 0
        self._keep = []

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().sr3_engine_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # ---- parameters
    def param_table(self):
        out = []
        n = lib().sr3_engine_num_params(self._h)
        buf = ctypes.create_string_buffer(256)
        shape = (c_int64 * 4)()
        nd = c_int()
        for i in range(n):
            _check(lib().sr3_engine_param_info(self._h, i, buf, 256, shape, ctypes.byref(nd)))
            out.append((buf.value.decode(), tuple(shape[j] for j in range(nd.value))))
        return out

    def load_params_fast(self, tensors):
        """One native call for the whole parameter set (fp32 contiguous CUDA tensors in param_table() order), no synchronisation: the
        training loop's re-pack after every optimizer step."""
        arr = (c_void_p * len(tensors))()
        for i, t in enumerate(tensors):
            if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
                raise RuntimeError("load_params_fast needs contiguous fp32 CUDA tensors")
            arr[i] = t.data_ptr()
        with torch.cuda.device(self.device):
            _check(lib().sr3_engine_load_all_params(self._h, arr, len(tensors), _stream()))

    def load_state_dict(self, sd: dict):
        keep = []
        with torch.cuda.device(self.device):
            for name, _shape in self.param_table():
                if name not in sd:
                    raise KeyError("missing key in state_dict: " + name)
                t = _f32c(sd[name], self.device)
                keep.append(t)
                _check(lib().sr3_engine_load_param(self._h, name.encode(), _ptr(t), t.numel(), _stream()))
            _check(lib().sr3_engine_finalize_params(self._h, _stream()))
            torch.cuda.current_stream().synchronize()

    def set_schedule(self, bufs: dict, sqrt_alphas_cumprod_prev):
        import numpy as np
        T = int(bufs["betas"].shape[0])
        host = [bufs[k].detach().to("cpu", torch.float32).contiguous() for k in
                ("sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod", "posterior_mean_coef1", "posterior_mean_coef2",
                 "posterior_log_variance_clipped")]
        sp = np.ascontiguousarray(np.asarray(sqrt_alphas_cumprod_prev, dtype=np.float64))
        assert sp.shape[0] == T + 1
        with torch.cuda.device(self.device):
            _check(lib().sr3_engine_set_schedule(self._h, T, *[c_void_p(h.data_ptr()) for h in host], c_void_p(sp.ctypes.data), _stream()))
        self.T = T

    # ---- compute
    def _img(self):
        return torch.empty(self.batch, self.channels, self.image_size, self.image_size, device=self.device, dtype=torch.float32)

    def unet_forward(self, x, noise_level):
        x = _f32c(x, self.device)
        nl = _f32c(noise_level, self.device).reshape(-1)
        assert x.shape == (self.batch, self.in_channel, self.image_size, self.image_size), x.shape
        assert nl.numel() == self.batch
        eps = torch.empty(self.batch, self.out_channel, self.image_size, self.image_size, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            _check(lib().sr3_unet_forward(self._h, _ptr(x), _ptr(nl), _ptr(eps), _stream()))
        return eps

    def p_mean_variance(self, x, t, clip_denoised=True, condition_x=None):
        x = _f32c(x, self.device)
        c = None if condition_x is None else _f32c(condition_x, self.device)
        mean = self._img()
        lv = c_float()
        with torch.cuda.device(self.device):
            _check(lib().sr3_p_mean_variance(self._h, _ptr(x), _ptr(c), int(t), int(bool(clip_denoised)), _ptr(mean), ctypes.byref(lv), _stream()))
        return mean, lv.value

    def p_sample(self, x, t, condition_x=None, noise=None, seed=0, first_index=0):
        x = _f32c(x, self.device)
        c = None if condition_x is None else _f32c(condition_x, self.device)
        n = None if noise is None else _f32c(noise, self.device)
        out = self._img()
        with torch.cuda.device(self.device):
            _check(lib().sr3_p_sample(self._h, _ptr(x), _ptr(c), int(t), _ptr(n), int(seed), int(first_index), _ptr(out), _stream()))
        return out

    def p_losses(self, hr, sr, gamma, noise, loss_type="l1"):
        hr, noise = _f32c(hr, self.device), _f32c(noise, self.device)
        s = None if sr is None else _f32c(sr, self.device)
        g = _f32c(gamma, self.device).reshape(-1)
        out = c_double()
        with torch.cuda.device(self.device):
            _check(lib().sr3_p_losses(self._h, _ptr(hr), _ptr(s), _ptr(g), _ptr(noise), 1 if loss_type == "l1" else 2, ctypes.byref(out), _stream()))
        return out.value

    # ---- training step (model/model.py:48-58): forward with the draws injected, backward into caller-owned gradient tensors
    def train_forward(self, hr, sr, gamma, noise, loss_type="l1", dropout_seed=0, want_loss=True):
        hr, noise = _f32c(hr, self.device), _f32c(noise, self.device)
        s = None if sr is None else _f32c(sr, self.device)
        g = _f32c(gamma, self.device).reshape(-1)
        self._keep = [hr, noise, s, g]
        out = c_double()
        with torch.cuda.device(self.device):
            _check(lib().sr3_train_forward(self._h, _ptr(hr), _ptr(s), _ptr(g), _ptr(noise), 1 if loss_type == "l1" else 2, int(dropout_seed),
                                           ctypes.byref(out) if want_loss else None, _stream()))
        return out.value if want_loss else None

    def _grad_ptrs(self, grads):
        arr = (c_void_p * len(grads))()
        for i, gten in enumerate(grads):
            assert gten.is_cuda and gten.dtype == torch.float32 and gten.is_contiguous()
            arr[i] = gten.data_ptr()
        return arr

    def train_backward(self, grad_scale, grads):
        """grads: one contiguous fp32 CUDA tensor per parameter, in param_table() order; overwritten."""
        arr = self._grad_ptrs(grads)
        with torch.cuda.device(self.device):
            _check(lib().sr3_train_backward(self._h, float(grad_scale), arr, len(grads), _stream()))

    def train_backward_profile(self, grad_scale, grads):
        """{kind: ms} of one backward, CUDA events around every op."""
        arr = self._grad_ptrs(grads)
        ms = (c_float * 8)()
        with torch.cuda.device(self.device):
            _check(lib().sr3_train_backward_profile(self._h, float(grad_scale), arr, len(grads), ms, _stream()))
        names = {0: "dgrad_tile_kernel", 1: "groupnorm_elementwise", 3: "bookkeeping", 4: "other", 5: "wgrad_slice_reduce", 6: "wgrad", 7: "attention_backward"}
        return {names.get(k, str(k)): ms[k] for k in range(8) if ms[k] > 0}

    def num_backward_blocks(self):
        return lib().sr3_train_num_backward_blocks(self._h)

    def backward_begin(self, grad_scale, grads):
        self._grad_arr = self._grad_ptrs(grads)
        _check(lib().sr3_train_backward_begin(self._h, float(grad_scale), self._grad_arr, len(grads)))

    def backward_block(self, i):
        with torch.cuda.device(self.device):
            _check(lib().sr3_train_backward_block(self._h, int(i), _stream()))

    def backward_flush(self):
        """Sum the weight-gradient partial tiles of the layers run since the last flush (one launch); call before all-reducing a bucket."""
        with torch.cuda.device(self.device):
            _check(lib().sr3_train_backward_flush(self._h, _stream()))

    def backward_finish(self):
        with torch.cuda.device(self.device):
            _check(lib().sr3_train_backward_finish(self._h, _stream()))

    def block_params(self, i):
        cap = 64
        idx = (c_int * cap)()
        n = c_int()
        _check(lib().sr3_train_block_params(self._h, int(i), idx, cap, ctypes.byref(n)))
        return [idx[k] for k in range(min(n.value, cap))]

    def dropout_layers(self):
        out = []
        buf = ctypes.create_string_buffer(256)
        for i in range(lib().sr3_train_num_dropout_layers(self._h)):
            _check(lib().sr3_train_dropout_layer_name(self._h, i, buf, 256))
            out.append(buf.value.decode())
        return out

    def set_dropout_mask(self, block_name, mask_nchw_u8):
        """Tests: inject the keep-mask (uint8 CUDA [B,C,H,W], 1 = keep) of the nn.Dropout of `block_name` ("downs.1.res_block.block2")."""
        if mask_nchw_u8 is not None:
            assert mask_nchw_u8.is_cuda and mask_nchw_u8.dtype == torch.uint8 and mask_nchw_u8.is_contiguous()
            self._keep_masks = getattr(self, "_keep_masks", {})
            self._keep_masks[block_name] = mask_nchw_u8
        _check(lib().sr3_train_set_dropout_mask(self._h, block_name.encode(), _ptr(mask_nchw_u8)))

    def p_sample_loop(self, condition_x, x_T, noises=None, seed=0, first_index=0, want_snapshots=True):
        c = None if condition_x is None else _f32c(condition_x, self.device)
        x_T = _f32c(x_T, self.device)
        n = None if noises is None else _f32c(noises, self.device)
        T = self.T
        inter = 1 | (T // 10)
        cap = len([i for i in range(T) if i % inter == 0])
        final = self._img()
        snaps = torch.empty(cap, *final.shape, device=self.device, dtype=torch.float32) if want_snapshots else None
        ns = c_int()
        with torch.cuda.device(self.device):
            _check(lib().sr3_p_sample_loop(self._h, _ptr(c), _ptr(x_T), _ptr(n), int(seed), int(first_index), _ptr(final), _ptr(snaps), cap,
                                           ctypes.byref(ns), _stream()))
        return final, snaps

    def super_resolution_host(self, cond_host, x_T_host, seed=0, first_index=0):
        """Host (pinned) buffers in, host buffer out; H2D + T steps + D2H inside one native call."""
        out = torch.empty(self.batch, self.channels, self.image_size, self.image_size, dtype=torch.float32).pin_memory()
        with torch.cuda.device(self.device):
            _check(lib().sr3_super_resolution_host(self._h, _ptr(cond_host), _ptr(x_T_host), int(seed), int(first_index), _ptr(out), _stream()))
        return out

    def loop_begin(self, condition_x, x_T, seed=0, first_index=0):
        c = None if condition_x is None else _f32c(condition_x, self.device)
        x_T = _f32c(x_T, self.device)
        with torch.cuda.device(self.device):
            _check(lib().sr3_p_sample_loop_begin(self._h, _ptr(c), _ptr(x_T), int(seed), int(first_index), _stream()))

    def steps(self, t_start, n):
        with torch.cuda.device(self.device):
            _check(lib().sr3_p_sample_steps(self._h, int(t_start), int(n), _stream()))

    def read_state(self):
        out = self._img()
        with torch.cuda.device(self.device):
            _check(lib().sr3_read_state(self._h, _ptr(out), _stream()))
        return out

    def profile_step(self, t, reps=3):
        """[(kind, ms, flops, bytes)] per launch of one eager step (kinds: 0 gemm tile, 1 GN apply, 2 cast, 3 softmax, 4 other)."""
        cap = lib().sr3_engine_num_ops_per_step(self._h)
        kinds, ms = (c_int * cap)(), (c_float * cap)()
        fl, by = (c_double * cap)(), (c_double * cap)()
        n = c_int()
        with torch.cuda.device(self.device):
            _check(lib().sr3_engine_profile_step(self._h, int(t), int(reps), cap, kinds, ms, fl, by, ctypes.byref(n), _stream()))
        return [(kinds[i], ms[i], fl[i], by[i]) for i in range(n.value)]

    def launches_per_step(self):
        return lib().sr3_engine_num_launches_per_step(self._h)

    def ops_per_step(self):
        return lib().sr3_engine_num_ops_per_step(self._h)

    def uses_step_kernel(self):
        return bool(lib().sr3_engine_uses_step_kernel(self._h))

    STEP_OP_NAMES = {0: "gemm_tile", 1: "groupnorm_apply", 2: "attention", 3: "softmax", 4: "embed_film", 5: "stats_clear"}

    def step_kernel_profile(self, phases=False):
        """[(op type, microseconds)] of the most recent persistent-step-kernel launch (device globaltimer stamps of CTA 0);
        phases=True: [(op type, us, (set-up, barrier wait, body, end fence))]."""
        cap = 4096
        types, us, ph = (c_int * cap)(), (c_double * cap)(), (c_double * (4 * cap))()
        n = c_int()
        with torch.cuda.device(self.device):
            _check(lib().sr3_engine_step_kernel_profile(self._h, cap, types, us, ph, ctypes.byref(n), _stream()))
        if phases:
            return [(types[i], us[i], tuple(ph[4 * i + k] for k in range(4))) for i in range(n.value)]
        return [(types[i], us[i]) for i in range(n.value)]

    def workspace_bytes(self):
        return lib().sr3_engine_workspace_bytes(self._h)

    def read_activation(self, name):
        """fp32 output of a top-level UNet layer of the last forward, returned NCHW like a reference forward hook."""
        numel = c_int64()
        shape = (c_int * 4)()
        _check(lib().sr3_engine_read_activation(self._h, name.encode(), c_void_p(0), 0, ctypes.byref(numel), shape, _stream()))
        t = torch.empty(tuple(shape), device=self.device, dtype=torch.float32)
        _check(lib().sr3_engine_read_activation(self._h, name.encode(), _ptr(t), t.numel(), ctypes.byref(numel), shape, _stream()))
        return t.permute(0, 3, 1, 2).contiguous()


def bench_conv(B, H, W, Cin, Cout, k=3, stride=1, resid=False, stats=True, reps=20):
    ms = c_float()
    _check(lib().sr3_bench_conv(B, H, W, Cin, Cout, k, stride, int(resid), int(stats), reps, ctypes.byref(ms)))
    return ms.value


def test_attention(qk_bf16, vT_bf16, nz, Lt, HW, C):
    """Fused attention core: qk [nz*Lt, 2C] bf16, vT [nz*C, Lt] bf16 -> O [nz*Lt, C] bf16."""
    out = torch.empty(nz * Lt, C, device=qk_bf16.device, dtype=torch.bfloat16)
    _check(lib().sr3_test_attention(_ptr(qk_bf16), _ptr(vT_bf16), _ptr(out), nz, Lt, HW, C, _stream()))
    return out


def test_gemm(a_bf16, b_bf16, block_n):
    M, K = a_bf16.shape
    N = b_bf16.shape[0]
    d = torch.empty(M, N, device=a_bf16.device, dtype=torch.float32)
    _check(lib().sr3_test_gemm(_ptr(a_bf16), _ptr(b_bf16), _ptr(d), M, N, K, block_n, _stream()))
    return d


def test_conv(x_nhwc_bf16, w_oihw, bias, ksize, stride, want_stats=False):
    B, H, W, Cin = x_nhwc_bf16.shape
    Cout = w_oihw.shape[0]
    y = torch.empty(B, H // stride, W // stride, Cout, device=x_nhwc_bf16.device, dtype=torch.float32)
    stats = torch.zeros(B, Cout, 2, device=y.device, dtype=torch.float64) if want_stats else None
    _check(lib().sr3_test_conv(_ptr(x_nhwc_bf16), _ptr(w_oihw), _ptr(bias), _ptr(y), _ptr(stats), B, H, W, Cin, Cout, ksize, stride, _stream()))
    return y, stats


def test_conv_groupnorm(x_nhwc_bf16, w_oihw, bias, gamma, beta, groups, silu, ksize):
    """conv (+ statistics in the epilogue) -> GroupNorm(+SiLU) apply: returns (y fp32 NHWC, a bf16 NHWC)."""
    B, H, W, Cin = x_nhwc_bf16.shape
    Cout = w_oihw.shape[0]
    y = torch.empty(B, H, W, Cout, device=x_nhwc_bf16.device, dtype=torch.float32)
    a = torch.empty(B, H, W, Cout, device=x_nhwc_bf16.device, dtype=torch.bfloat16)
    _check(lib().sr3_test_conv_groupnorm(_ptr(x_nhwc_bf16), _ptr(w_oihw), _ptr(bias), _ptr(gamma), _ptr(beta), groups, int(silu), _ptr(y), _ptr(a),
                                         B, H, W, Cin, Cout, ksize, _stream()))
    return y, a


def adam_step(table_dev, n_tensors, lr, beta1, beta2, eps, step, grad_scale=1.0):
    """torch.optim.Adam step over a device table of {param, grad, exp_avg, exp_avg_sq, numel} records (int64 [n, 5]) in one launch."""
    _check(lib().sr3_adam_step(_ptr(table_dev), int(n_tensors), float(lr), float(beta1), float(beta2), float(eps), int(step), float(grad_scale), _stream()))
