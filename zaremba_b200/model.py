"""Drop-in replacement for the reference's `model.py` (`from model import Model`).

Same constructor arguments, parameter names / shapes / registration order, RNG
consumption at construction, `(h, c)` state layouts and `forward(x, states)` contract as
/root/reference/model.py:75-110 -- but every tensor operation of the forward and backward
pass runs in libzaremba_b200.so (hand-written sm_100a CUDA behind the C ABI of
include/zaremba_b200.h).  PyTorch is used for device memory, streams and autograd glue
only.  There is no CPU path: the module refuses to run off a CUDA device.

Reference lines mirrored:
  Embed / LSTM / Linear containers   model.py:6-71   (parameter holders here; names kept)
  Model.__init__ / reset_parameters  model.py:76-92
  state_init / detach                model.py:94-101
  forward                            model.py:103-110
"""
from __future__ import annotations

import ctypes as C
import math

import torch
from torch import nn

from . import _lib


class Embed(nn.Module):
    """Parameter holder for `embed.W` [V,H] (model.py:6-17)."""

    def __init__(self, vocab_size, embed_size):
        super().__init__()
        self.vocab_size, self.embed_size = vocab_size, embed_size
        self.W = nn.Parameter(torch.empty(vocab_size, embed_size))

    def extra_repr(self):
        return f"vocab: {self.vocab_size}, embedding: {self.embed_size}"


class LSTM(nn.Module):
    """Parameter holder for one recurrent layer.

    lstm_type "pytorch": torch.nn.LSTM names and gate order (i,f,g,o) (model.py:84);
    constructing it consumes the global RNG exactly like nn.LSTM.__init__ does (four
    U(-1/sqrt(H), 1/sqrt(H)) draws), so that a seeded `Model(...)` gets the reference's
    weights.  lstm_type "custom": the reference's own cell (model.py:20-31): names
    W_x/W_h/b_x/b_h, row blocks (i,f,o,n), no RNG consumption.
    """

    def __init__(self, input_size, hidden_size, lstm_type="pytorch"):
        super().__init__()
        assert input_size == hidden_size, "the reference only builds H->H layers"
        self.input_size, self.hidden_size, self.lstm_type = input_size, hidden_size, lstm_type
        H = hidden_size
        if lstm_type == "custom":
            self.W_x = nn.Parameter(torch.empty(4 * H, H))
            self.W_h = nn.Parameter(torch.empty(4 * H, H))
            self.b_x = nn.Parameter(torch.empty(4 * H))
            self.b_h = nn.Parameter(torch.empty(4 * H))
        else:
            stdv = 1.0 / math.sqrt(H)
            self.weight_ih_l0 = nn.Parameter(torch.empty(4 * H, H).uniform_(-stdv, stdv))
            self.weight_hh_l0 = nn.Parameter(torch.empty(4 * H, H).uniform_(-stdv, stdv))
            self.bias_ih_l0 = nn.Parameter(torch.empty(4 * H).uniform_(-stdv, stdv))
            self.bias_hh_l0 = nn.Parameter(torch.empty(4 * H).uniform_(-stdv, stdv))

    def tensors(self):
        if self.lstm_type == "custom":
            return self.W_x, self.W_h, self.b_x, self.b_h
        return self.weight_ih_l0, self.weight_hh_l0, self.bias_ih_l0, self.bias_hh_l0

    def extra_repr(self):
        return f"input: {self.input_size}, hidden: {self.hidden_size}, type: {self.lstm_type}"


class Linear(nn.Module):
    """Parameter holder for `fc.W` [V,H], `fc.b` [V] (model.py:57-71)."""

    def __init__(self, input_size, hidden_size):
        super().__init__()
        self.input_size, self.hidden_size = input_size, hidden_size
        self.W = nn.Parameter(torch.empty(hidden_size, input_size))
        self.b = nn.Parameter(torch.empty(hidden_size))

    def extra_repr(self):
        return f"input: {self.input_size}, output: {self.hidden_size}"


def _ifon_to_ifgo(t):
    """custom cell row blocks (i,f,o,n) <-> nn.LSTM (i,f,g,o); the permutation is an involution."""
    i, f, a, b = t.chunk(4, 0)
    return torch.cat([i, f, b, a], 0)


class _LmFunction(torch.autograd.Function):
    """autograd node for model.py:103-110: forward = zrb_forward, backward = zrb_backward."""

    @staticmethod
    def forward(ctx, model, x_dev, states_in, seed, step, *weights):
        scores, states_out = model._run_forward(x_dev, states_in, weights, seed, step, want_scores=True)
        ctx.model = model
        ctx.fwd_id = model._fwd_id
        ctx.save_for_backward(*weights)
        flat = [t for hc in states_out for t in hc]
        ctx.mark_non_differentiable(*flat)
        return (scores, *flat)

    @staticmethod
    def backward(ctx, dscores, *unused):
        model = ctx.model
        if ctx.fwd_id != model._fwd_id:
            raise RuntimeError("zaremba_b200.Model keeps activations of the latest forward only; "
                               "backward() must follow the forward it belongs to")
        grads = model._run_backward(dscores.contiguous(), ctx.saved_tensors)
        return (None, None, None, None, None, *grads)


class Model(nn.Module):
    """`Model(vocab_size, hidden_size, layer_num, dropout, winit, lstm_type="pytorch")`.

    Extra keyword `engine`: "tc" (tcgen05 tensor cores, default) or "simt" (fp32 CUDA cores,
    validation).
    """

    def __init__(self, vocab_size, hidden_size, layer_num, dropout, winit, lstm_type="pytorch", engine="tc"):
        super().__init__()
        if lstm_type not in ("pytorch", "custom"):
            raise ValueError(f"lstm_type must be 'pytorch' or 'custom', got {lstm_type!r}")
        if layer_num > _lib.MAX_LAYERS:
            raise ValueError(f"at most {_lib.MAX_LAYERS} layers")
        self.vocab_size = vocab_size
        self.hidden_size = hidden_size
        self.layer_num = layer_num
        self.winit = winit
        self.lstm_type = lstm_type
        self.engine = engine
        self.p_drop = float(dropout)
        self.embed = Embed(vocab_size, hidden_size)
        self.rnns = nn.ModuleList(LSTM(hidden_size, hidden_size, lstm_type) for _ in range(layer_num))
        self.fc = Linear(hidden_size, vocab_size)
        self.dropout = nn.Dropout(p=dropout)     # kept for repr / state parity; masks come from the library
        self.reset_parameters()
        self._ctx = None
        self._ctx_key = None
        self._fwd_id = 0
        self._drop_step = 0
        self._seed = None
        self._versions = None
        self._explicit_masks = None

    # ---- reference API -----------------------------------------------------------------
    def reset_parameters(self):
        for param in self.parameters():          # model.py:90-92
            nn.init.uniform_(param, -self.winit, self.winit)

    def state_init(self, batch_size):
        dev = next(self.parameters()).device
        shape = (batch_size, self.hidden_size) if self.lstm_type == "custom" else (1, batch_size, self.hidden_size)
        return [(torch.zeros(shape, device=dev), torch.zeros(shape, device=dev)) for _ in self.rnns]

    def detach(self, states):
        return [(h.detach(), c.detach()) for (h, c) in states]

    def forward(self, x, states):
        dev = self.embed.W.device
        if dev.type != "cuda":
            raise RuntimeError("zaremba_b200.Model runs on a CUDA device only (no CPU fallback): call .to('cuda')")
        x_dev = x.to(device=dev, dtype=torch.int64).contiguous()   # main.py hands CPU non-contiguous views
        weights = self._lib_weights()
        seed, step = self._next_dropout_key()
        need_grad = torch.is_grad_enabled() and any(w.requires_grad for w in weights)
        if need_grad:
            outs = _LmFunction.apply(self, x_dev, states, seed, step, *weights)
            scores, flat = outs[0], outs[1:]
            new_states = [(flat[2 * i], flat[2 * i + 1]) for i in range(self.layer_num)]
        else:
            scores, new_states = self._run_forward(x_dev, states, weights, seed, step, want_scores=True)
        for i in range(self.layer_num):          # the reference mutates the caller's list (model.py:107)
            states[i] = new_states[i]
        return scores, states

    # ---- plumbing ------------------------------------------------------------------------
    def ordered_parameters(self):
        """The 3+4L tensors in registration order, as the library's zrb_params expects them
        (pytorch names / gate order; the custom layout is permuted by `_lib_weights`)."""
        out = [self.embed.W]
        for r in self.rnns:
            out += list(r.tensors())
        out += [self.fc.W, self.fc.b]
        return out

    def _lib_weights(self):
        ws = self.ordered_parameters()
        if self.lstm_type == "custom":
            # (i,f,o,n) -> (i,f,g,o) row-block permutation: a differentiable copy, so autograd
            # routes the gradients back into the custom layout
            ws = [w if i in (0, len(ws) - 2, len(ws) - 1) else _ifon_to_ifgo(w) for i, w in enumerate(ws)]
        return ws

    def _next_dropout_key(self):
        if self._seed is None:
            self._seed = int(torch.initial_seed()) & 0xFFFFFFFFFFFFFFFF
        step = self._drop_step
        if self.training:
            self._drop_step += 1
        return self._seed, step

    def set_explicit_dropout_masks(self, masks):
        """Replay given keep-masks (list of L+1 uint8/bool CUDA tensors [T,B,H]) instead of
        Philox; None restores Philox.  Used by parity tests with the reference's masks."""
        self._explicit_masks = None if masks is None else [m.to(torch.uint8).contiguous() for m in masks]
        if self._ctx is not None:
            self._push_masks()

    def _push_masks(self):
        lib = _lib.load()
        if self._explicit_masks is None:
            _lib.check(lib.zrb_set_explicit_masks(self._ctx, None))
        else:
            arr = (C.c_void_p * (self.layer_num + 1))(*[m.data_ptr() for m in self._explicit_masks])
            _lib.check(lib.zrb_set_explicit_masks(self._ctx, arr))

    def _context(self, T, B):
        key = (max(T, 1), max(B, 1), self.embed.W.device.index)
        if self._ctx is not None:
            ok = self._ctx_key[2] == key[2] and self._ctx_key[0] >= T and self._ctx_key[1] >= B
            if ok:
                return self._ctx
            self._destroy_ctx()
        lib = _lib.load()
        cfg = _lib.ZrbConfig(self.vocab_size, self.hidden_size, self.layer_num, key[0], key[1],
                             _lib.ENGINE_TC if self.engine == "tc" else _lib.ENGINE_SIMT, self.p_drop, 0)
        h = C.c_void_p()
        with torch.cuda.device(self.embed.W.device):
            _lib.check(lib.zrb_ctx_create(C.byref(cfg), C.byref(h)))
        self._ctx, self._ctx_key = h, key
        self._versions = None
        if self._explicit_masks is not None:
            self._push_masks()
        return self._ctx

    def _destroy_ctx(self):
        if getattr(self, "_ctx", None) is not None:
            _lib.load().zrb_ctx_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self._destroy_ctx()
        except Exception:
            pass

    def _params_struct(self, tensors):
        L = self.layer_num
        ps = _lib.ZrbParams()
        for t in tensors:
            if t.dtype != torch.float32 or not t.is_cuda:
                raise RuntimeError("parameters must be fp32 CUDA tensors")
        ts = [t if t.is_contiguous() else t.contiguous() for t in tensors]
        ps.embed_w = ts[0].data_ptr()
        for l in range(L):
            ps.w_ih[l] = ts[1 + 4 * l].data_ptr()
            ps.w_hh[l] = ts[2 + 4 * l].data_ptr()
            ps.b_ih[l] = ts[3 + 4 * l].data_ptr()
            ps.b_hh[l] = ts[4 + 4 * l].data_ptr()
        ps.fc_w = ts[1 + 4 * L].data_ptr()
        ps.fc_b = ts[2 + 4 * L].data_ptr()
        return ps, ts

    def _states_struct(self, states):
        st = _lib.ZrbStates()
        keep = []
        for l, (h, c) in enumerate(states):
            h = h.detach().to(torch.float32).contiguous()
            c = c.detach().to(torch.float32).contiguous()
            keep += [h, c]
            st.h[l] = h.data_ptr()
            st.c[l] = c.data_ptr()
        return st, keep

    def _note_param_versions(self):
        """main.py:116-117 updates parameters in place outside the library: tell the context so
        it rebuilds its low-precision weight images."""
        v = tuple((p.data_ptr(), p._version) for p in self.parameters())
        if v != self._versions:
            _lib.check(_lib.load().zrb_params_changed(self._ctx))
            self._versions = v

    def _run_forward(self, x_dev, states, weights, seed, step, want_scores=True):
        lib = _lib.load()
        T, B = x_dev.shape
        ctx = self._context(T, B)
        self._note_param_versions()
        dev = x_dev.device
        ps, keep_w = self._params_struct([w.detach() for w in weights])
        st_in, keep_in = self._states_struct(states)
        out_states = [(torch.empty_like(h, dtype=torch.float32), torch.empty_like(c, dtype=torch.float32))
                      for (h, c) in states]
        st_out, keep_out = self._states_struct(out_states)
        scores = torch.empty(T * B, self.vocab_size, device=dev, dtype=torch.float32) if want_scores else None
        stream = torch.cuda.current_stream(dev).cuda_stream
        with torch.cuda.device(dev):
            _lib.check(lib.zrb_forward(ctx, C.byref(ps), _lib.ptr(x_dev), T, B, C.byref(st_in), C.byref(st_out),
                                       _lib.ptr(scores), 1 if self.training else 0, seed, step, stream))
        # _states_struct made contiguous detached aliases of out_states' storage
        self._fwd_id += 1
        return scores, out_states

    def _run_backward(self, dscores, weights):
        lib = _lib.load()
        dev = dscores.device
        ps, keep_w = self._params_struct([w.detach() for w in weights])
        grads = [torch.empty_like(w) for w in weights]
        gs, keep_g = self._params_struct(grads)
        stream = torch.cuda.current_stream(dev).cuda_stream
        with torch.cuda.device(dev):
            _lib.check(lib.zrb_backward(self._ctx, C.byref(ps), _lib.ptr(dscores.to(torch.float32)), C.byref(gs),
                                        stream))
        return grads
