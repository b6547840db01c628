"""B200-native drop-in for the inference surface of the reference's ``models/voicecraft.py``.

Same constructor (``VoiceCraft(args)`` / ``VoiceCraft(config=dict)``), same ``state_dict`` keys, same
``inference_tts`` / ``inference_tts_batch`` / ``inference`` signatures and return shapes
(reference models/voicecraft.py:97-121, 561-573, 908-920, 1156-1169).  The module only *holds* the
parameters; every inference call goes through libvcb200.so (hand-written sm_100a kernels: paged-KV
attention, tcgen05 GEMMs, fused sampler).  There is no PyTorch / CPU fallback: without the extension or
without a Blackwell GPU the calls raise.

Random numbers: the reference samples with ``torch.multinomial(softmax(l), 1)``, which ATen evaluates as
``argmax(softmax(l) / q)`` with ``q = empty_like(p).exponential_(1)`` from the device's global generator.
The fused sampler kernel generates exactly that ``q`` itself: every utterance (or best-of-N group) owns the
Philox stream of a torch CUDA generator at (seed, offset) and consumes, per sampling step, what the reference's
draw of shape ``[n*K, V]`` consumes.  ``inference_tts`` / ``inference_tts_batch`` / ``inference`` take the
stream of the model device's default generator and leave it advanced as the reference would; batched sessions
give every utterance its own seed, so row *i* of a batch equals the single call of utterance *i* under that
seed.  ``noise_fn`` replaces the generator (tests feed CPU-generator noise to compare with the CPU oracle).

Out of scope (training): ``forward`` and ``prepare_mask_intervals`` raise NotImplementedError.
"""
import copy
import ctypes as C
import logging
import os
from argparse import Namespace
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .codebooks_patterns import DelayedPatternProvider

try:  # same mixin as the reference (voicecraft.py:23, 89-95); optional so the hot path has no hard dependency
    from huggingface_hub import PyTorchModelHubMixin
    _HubBase = (PyTorchModelHubMixin,)
    _HUB_KW = dict(library_name="voicecraft", repo_url="https://github.com/jasonppy/VoiceCraft", tags=["text-to-speech"])
except Exception:  # pragma: no cover
    _HubBase = ()
    _HUB_KW = {}


def sine_pe(length: int, dim: int) -> torch.Tensor:
    """Sinusoidal table of SinePositionalEmbedding.extend_pe (embedding.py:67-92), fp32 [length, dim]."""
    import math
    pe = torch.zeros(length, dim)
    position = torch.arange(0, length, dtype=torch.float32).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, dim, 2, dtype=torch.float32) * -(math.log(10000.0) / dim))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe


# ---------------------------------------------------------------------------------------------------------
# parameter containers with the reference's attribute names (so state_dict keys match)
# ---------------------------------------------------------------------------------------------------------
class _TokenEmbedding(nn.Module):
    def __init__(self, dim, vocab):
        super().__init__()
        self.word_embeddings = nn.Embedding(vocab, dim)


class _Alpha(nn.Module):
    def __init__(self):
        super().__init__()
        self.alpha = nn.Parameter(torch.ones(1))


class _OutProj(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(d, d))
        self.bias = nn.Parameter(torch.zeros(d))
        nn.init.xavier_uniform_(self.weight)


class _SelfAttn(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.in_proj_weight = nn.Parameter(torch.empty(3 * d, d))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * d))
        self.out_proj = _OutProj(d)
        nn.init.xavier_uniform_(self.in_proj_weight)


class _Layer(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.self_attn = _SelfAttn(d)
        self.linear1 = nn.Linear(d, 4 * d)
        self.linear2 = nn.Linear(4 * d, d)
        self.norm1 = nn.LayerNorm(d, eps=1e-5)
        self.norm2 = nn.LayerNorm(d, eps=1e-5)


class _Decoder(nn.Module):
    def __init__(self, d, n_layers):
        super().__init__()
        self.layers = nn.ModuleList([_Layer(d) for _ in range(n_layers)])
        self.norm = nn.LayerNorm(d, eps=1e-5)


class VoiceCraft(nn.Module, *_HubBase, **_HUB_KW):
    def __new__(cls, args: Optional[Namespace] = None, config: Optional[Dict] = None, **kwargs):
        if args is not None:
            if config is not None:
                raise ValueError("Cannot provide both `args` and `config`.")
            config = vars(args)
        if _HubBase:
            return super().__new__(cls, args=args, config=config, **kwargs)
        return super().__new__(cls)

    def __init__(self, args: Optional[Namespace] = None, config: Optional[Dict] = None):
        super().__init__()
        if args is None:
            if config is None:
                raise ValueError("Either `args` or `config` must be provided.")
            args = Namespace(**config)
        a = self.args = copy.copy(args)
        self.pattern = DelayedPatternProvider(n_q=a.n_codebooks)
        if not getattr(a, "special_first", False):
            a.special_first = 0
        if not getattr(a, "n_special", False):
            a.n_special = 3
        a.eos = getattr(a, "eos", -1)
        K, d = a.n_codebooks, a.d_model
        self.eog = nn.Parameter(torch.full((K, 1), a.eog, dtype=torch.long), requires_grad=False)
        if a.eos > 0:
            assert a.eos != a.audio_pad_token and a.eos != a.empty_token, a.eos
            self.eos = nn.Parameter(torch.full((K, 1), a.eos, dtype=torch.long), requires_grad=False)
        if isinstance(a.audio_vocab_size, str):
            a.audio_vocab_size = eval(a.audio_vocab_size)
        self.n_text_tokens = a.text_vocab_size + 1
        assert a.text_pad_token == a.text_vocab_size
        self.n_audio_tokens = [a.audio_vocab_size + a.n_special] * K
        assert a.audio_vocab_size == a.empty_token, a.empty_token
        assert a.eog == a.audio_vocab_size + 1, a.eog
        assert a.audio_pad_token == a.audio_vocab_size + 2, a.audio_pad_token
        assert getattr(a, "audio_embedding_dim", d) == d, "audio_embedding_dim must equal d_model (summed embeddings)"

        self.text_embedding = _TokenEmbedding(d, self.n_text_tokens)
        self.audio_embedding = nn.ModuleList([_TokenEmbedding(d, self.n_audio_tokens[k]) for k in range(K)])
        self.mask_embedding = nn.Parameter(torch.randn(a.max_n_spans, d), requires_grad=True)
        self.text_positional_embedding = _Alpha()
        self.audio_positional_embedding = _Alpha()
        self.decoder = _Decoder(d, a.num_decoder_layers)
        self.predict_layer = nn.ModuleList([
            nn.Sequential(nn.Linear(d, a.audio_vocab_size // 2), nn.GELU(),
                          nn.Linear(a.audio_vocab_size // 2, self.n_audio_tokens[k])) for k in range(K)])

        # engine configuration (see configure_engine)
        self._eng = None
        self._eng_key = None
        self._eng_opts = dict(max_slots=8, max_seq_len=2048, max_new_tokens=4096, kv_dtype="bf16")
        self.noise_fn = None          # optional: callable(shape, device) -> fp32 Exp(1) tensor on `device`
        self.poll_every = 4           # inference_tts*: poll the done flag every N steps (device generator only)
        self._sessions = set()        # open DecodeSessions (they hold engine slots; the engine is not rebuilt under them)
        self.last_stats = {}
        self.trace_logits = None      # set to a list to collect the raw logits [n*K, V] of every sampling step

    # ------------------------------------------------------------------------------------------------
    # training surface: out of scope for this build
    # ------------------------------------------------------------------------------------------------
    def forward(self, batch):
        raise NotImplementedError("training forward is out of scope of the B200 decode engine "
                                  "(reference models/voicecraft.py:472-559)")

    def prepare_mask_intervals(self, y_lens):
        raise NotImplementedError("training-only helper (reference models/voicecraft.py:198-237)")

    # ------------------------------------------------------------------------------------------------
    # engine management
    # ------------------------------------------------------------------------------------------------
    def configure_engine(self, **opts):
        """max_slots, max_seq_len, max_new_tokens, kv_dtype ('bf16' default | 'fp32').  Rebuilds lazily."""
        for k in opts:
            if k not in self._eng_opts:
                raise KeyError(k)
        self._eng_opts.update(opts)
        self._drop_engine()

    def _drop_engine(self):
        if getattr(self, "_eng", None) is not None:
            live = [s for s in getattr(self, "_sessions", ()) if s._open]
            if live:
                raise _lib.VcbError(f"{len(live)} DecodeSession(s) still hold slots of this engine: close them before "
                                    "reconfiguring / moving / reloading the model")
            _lib.load().vcb_destroy(self._eng)
        self._eng = None
        self._eng_key = None

    def _free_slots(self, n, eng_slots):
        """first of n consecutive engine slots not held by an open session (single calls and sessions share one engine)"""
        used = set()
        for s in self._sessions:
            if s._open:
                used.update(s.slots)
        for base in range(0, eng_slots - n + 1):
            if not any((base + i) in used for i in range(n)):
                return base
        raise _lib.VcbError("no free engine slots")

    def __del__(self):
        try:
            for sess in list(getattr(self, "_sessions", ())):
                sess.close()
            self._drop_engine()
        except Exception:
            pass

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        state_dict = {k: v for k, v in state_dict.items() if not k.startswith("accuracy_metrics")}
        out = super().load_state_dict(state_dict, strict=strict, **kw)
        self._drop_engine()
        return out

    def _apply(self, fn, *a, **kw):
        out = super()._apply(fn, *a, **kw)
        if hasattr(self, "_eng"):
            self._drop_engine()
        return out

    def _engine(self, need_slots=1, need_seq=0):
        dev = self.mask_embedding.device
        if dev.type != "cuda":
            raise _lib.VcbError("VoiceCraft (B200) has no CPU path: move the model to a CUDA device (`.to('cuda')`)")
        o = self._eng_opts
        held = sum(len(s.slots) for s in self._sessions if s._open)
        need_slots += held
        if need_slots > o["max_slots"] or need_seq > o["max_seq_len"]:
            if held:
                raise _lib.VcbError(f"engine too small (max_slots={o['max_slots']}, max_seq_len={o['max_seq_len']}) and "
                                    f"{held} slot(s) are held by open DecodeSessions: close them or configure_engine() first")
            o["max_slots"] = max(o["max_slots"], need_slots)
            o["max_seq_len"] = max(o["max_seq_len"], (need_seq + 255) // 256 * 256)
            self._drop_engine()
        key = (dev.index or 0, tuple(sorted(o.items())))
        if self._eng is not None and self._eng_key == key:
            return self._eng
        self._drop_engine()
        lib = _lib.load()
        a = self.args
        cfg = _lib.vcb_config(
            d_model=a.d_model, nhead=a.nhead, num_layers=a.num_decoder_layers, n_codebooks=a.n_codebooks,
            audio_vocab_size=a.audio_vocab_size, n_special=a.n_special, text_vocab_rows=self.n_text_tokens,
            empty_token=a.empty_token, eog=a.eog, audio_pad_token=a.audio_pad_token, eos=a.eos if a.eos > 0 else -1,
            encodec_sr=int(a.encodec_sr), max_n_spans=a.max_n_spans, max_slots=o["max_slots"],
            max_seq_len=o["max_seq_len"], max_new_tokens=o["max_new_tokens"],
            kv_dtype=1 if o["kv_dtype"] == "fp32" else 0, device=dev.index or 0)
        h = C.c_void_p()
        _lib.check(lib.vcb_create(C.byref(cfg), C.byref(h)))
        try:
            with torch.cuda.device(dev):
                for k, v in self.state_dict().items():
                    if not v.is_floating_point():
                        continue
                    t = v.detach().to(device=dev, dtype=torch.float32).contiguous()
                    shape = (C.c_int64 * max(t.dim(), 1))(*t.shape) if t.dim() else (C.c_int64 * 1)(1)
                    _lib.check(lib.vcb_load_weight(h, k.encode(), t.data_ptr(), shape, max(t.dim(), 1), 1))
                pe = sine_pe(max(4000, o["max_seq_len"]), a.d_model).to(dev)
                _lib.check(lib.vcb_load_pe(h, pe.data_ptr(), pe.shape[0], 1))
                _lib.check(lib.vcb_finalize_weights(h))
        except Exception:
            lib.vcb_destroy(h)
            raise
        self._eng, self._eng_key = h, key
        return h

    # ------------------------------------------------------------------------------------------------
    # helpers
    # ------------------------------------------------------------------------------------------------
    def _sampling(self, top_k, top_p, temperature, stop_repetition, silence_tokens):
        sp = _lib.vcb_sampling(top_k=int(top_k), top_p=float(top_p), temperature=float(temperature),
                               stop_repetition=int(stop_repetition), n_silence=min(len(silence_tokens), 8))
        for i, t in enumerate(list(silence_tokens)[:8]):
            sp.silence_tokens[i] = int(t)
        return sp

    @staticmethod
    def _rng_threads(dev, numel):
        """threads of ATen's distribution_nullary_kernel for a draw of `numel` elements (calc_execution_policy)"""
        p = torch.cuda.get_device_properties(dev)
        grid = min((numel + 255) // 256, p.multi_processor_count * (p.max_threads_per_multi_processor // 256))
        return 256 * grid

    def _check_ids(self, x_ids, y_tok):
        """the reference raises on an out-of-range id (nn.Embedding / F.embedding); the kernels index raw tables"""
        if x_ids.numel() and (int(x_ids.min()) < 0 or int(x_ids.max()) >= self.n_text_tokens):
            raise IndexError(f"text id out of range [0, {self.n_text_tokens})")
        if y_tok.numel() and (int(y_tok.min()) < 0 or int(y_tok.max()) >= self.n_audio_tokens[0]):
            raise IndexError(f"audio token out of range [0, {self.n_audio_tokens[0]})")

    def _draw_noise(self, buf):
        """caller-provided Exp(1) noise with the call shape of the reference's multinomial draw"""
        q = self.noise_fn(tuple(buf.shape), buf.device)
        buf.copy_(q.to(device=buf.device, dtype=torch.float32))
        return buf

    def shift(self, rearranged_y):
        """Delay every segment with the codebook pattern (reference voicecraft.py:254-262)."""
        shifted_y, patterns = [], []
        for segs in rearranged_y:
            pats = [self.pattern.get_pattern(s.shape[1]) for s in segs]
            out = [p.build_pattern_sequence(z=s.unsqueeze(0).contiguous(), special_token=self.args.empty_token,
                                            keep_only_valid_steps=False) for p, s in zip(pats, segs)]
            shifted_y.append([o[0].squeeze(0) for o in out])
            patterns.append(pats)
        return shifted_y, patterns

    def _run(self, eng, slots, n_rows, sp, stream, max_steps=None, speculative=False):
        """prefill is done; run sample / decode_step until every listed slot's group is done.

        Device generator (noise_fn is None): the sampler draws from the group's own Philox stream; a finished group
        ignores further steps and consumes nothing, so the done flag is polled only every `poll_every` steps
        (speculative, TTS) and the generator still ends exactly where the reference's would."""
        lib = _lib.load()
        a = self.args
        dev = self.mask_embedding.device
        K, V = a.n_codebooks, self.n_audio_tokens[0]
        n = len(slots)
        c_slots = (C.c_int32 * n)(*slots)
        host_noise = self.noise_fn is not None
        buf = torch.empty((n_rows * K, V), device=dev, dtype=torch.float32) if host_noise else None
        status = (_lib.vcb_status * n)()
        def trace():
            if self.trace_logits is not None:
                t = torch.empty(n_rows * K, V, device=dev, dtype=torch.float32)
                _lib.check(lib.vcb_debug_logits(eng, t.data_ptr(), n_rows * K))
                self.trace_logits.append(t)
        every = max(1, int(self.poll_every)) if (speculative and not host_noise and self.trace_logits is None) else 1
        noise = self._draw_noise(buf).data_ptr() if host_noise else None
        _lib.check(lib.vcb_sample(eng, c_slots, n, noise, C.byref(sp), stream))
        trace()
        steps = 1
        while True:
            if steps % every == 0 or every == 1:
                _lib.check(lib.vcb_poll(eng, c_slots, n, status, stream))
                if any(s.done == 2 for s in status):
                    raise _lib.VcbError("decode stopped: engine capacity (max_new_tokens / max_seq_len) exhausted; "
                                        "raise it with configure_engine()")
                if all(s.done for s in status):
                    break
            if max_steps is not None and steps >= max_steps:
                break
            forced = every == 1 and any(s.forced for s in status)
            if host_noise and not forced:                # forced hand-over steps consume no random numbers
                noise = self._draw_noise(buf).data_ptr()
            _lib.check(lib.vcb_decode_step(eng, c_slots, n, noise, C.byref(sp), stream))
            if not forced:
                trace()
            steps += 1
        return status

    def _device_rng(self, P, dev, n_rows):
        """hand the model device's default generator stream to the prompt (no-op with a caller noise_fn)"""
        if self.noise_fn is not None:
            return None
        gen = torch.cuda.default_generators[dev.index or 0]
        P.rng_seed = int(gen.initial_seed()) & 0xFFFFFFFFFFFFFFFF
        P.rng_offset = int(gen.get_offset())
        P.rng_threads = self._rng_threads(dev, n_rows * self.args.n_codebooks * self.n_audio_tokens[0])
        return gen

    def _read_rows(self, eng, slot, n_steps, stream):
        K = self.args.n_codebooks
        buf = (C.c_int32 * (n_steps * K))()
        _lib.check(_lib.load().vcb_read_tokens(eng, slot, buf, n_steps, stream))
        return np.frombuffer(buf, dtype=np.int32).reshape(n_steps, K).astype(np.int64)

    @staticmethod
    def _undelay(rows: np.ndarray, K: int) -> np.ndarray:
        """rows [n,K] (delayed, as sampled) -> [K, n-K]   (reference voicecraft.py:1126-1137)."""
        n = rows.shape[0]
        return np.stack([rows[k: n - (K - k), k] for k in range(K)], axis=0)

    # ------------------------------------------------------------------------------------------------
    # inference_tts  (reference voicecraft.py:908-1153)
    # ------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def inference_tts(self, x: torch.Tensor, x_lens: torch.Tensor, y: torch.Tensor, top_k: int = -100,
                      top_p: float = 1.0, temperature: float = 1.0, stop_repetition: int = 3, kvcache: int = 1,
                      silence_tokens: List[int] = [1388, 1898, 131], *kargs):
        res, gen = self._tts_impl(x, x_lens, y, top_k, top_p, temperature, stop_repetition, silence_tokens, 1)
        return res, gen

    @torch.no_grad()
    def inference_tts_batch(self, x: torch.Tensor, x_lens: torch.Tensor, y: torch.Tensor, top_k: int = -100,
                            top_p: float = 1.0, temperature: float = 1.0, stop_repetition: int = 3, kvcache: int = 1,
                            batch_size: int = 5, silence_tokens: List[int] = [1388, 1898, 131], *kargs):
        """Best-of-N: the first sample to end wins (reference voicecraft.py:1156-1439)."""
        return self._tts_impl(x, x_lens, y, top_k, top_p, temperature, stop_repetition, silence_tokens, batch_size)

    def _tts_impl(self, x, x_lens, y, top_k, top_p, temperature, stop_repetition, silence_tokens, n_copies):
        a = self.args
        K = a.n_codebooks
        assert x.ndim == 2, x.shape
        assert x_lens.ndim == 1, x_lens.shape
        assert y.ndim == 3, y.shape
        dev = self.mask_embedding.device
        x = x.to(dev)
        y = y.to(dev)
        if a.special_first:
            y = y + int(a.n_special)
        y = y.transpose(2, 1)                                     # [1,T,K] -> [1,K,T]
        assert y.shape[0] == 1 and y.shape[1] == K, y.shape
        logging.info(f"silence tokens: {silence_tokens}, note that if you are not using the pretrained encodec "
                     f"6f79c6a8, make sure you specified it yourself, rather than using the default")
        shifted, _ = self.shift([[y[0].long()]])
        prompt = shifted[0][0][:, : -(K - 1)] if K > 1 else shifted[0][0]     # voicecraft.py:967
        y_tok = prompt.transpose(1, 0).contiguous()                           # [T+1, K]
        x_len, y_len = int(x.shape[1]), int(y_tok.shape[0])
        cap = x_len * (int(a.encodec_sr) // 5)
        need_seq = x_len + max(y_len, cap + 1) + K + 8
        eng = self._engine(need_slots=n_copies, need_seq=need_seq)
        lib = _lib.load()
        x_ids = x[0].long().contiguous()
        self._check_ids(x_ids, y_tok)
        sp = self._sampling(top_k, top_p, temperature, stop_repetition, silence_tokens)
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream().cuda_stream
            base = self._free_slots(n_copies, self._eng_opts["max_slots"])
            P = _lib.vcb_prompt(slot=base, n_copies=n_copies, mode=0, x_len=x_len, text_ids_dev=x_ids.data_ptr(),
                                y_len=y_len, y_tokens_dev=y_tok.data_ptr(), mask_rows_dev=None, n_more_spans=0)
            gen_state = self._device_rng(P, dev, n_copies)
            _lib.check(lib.vcb_prefill(eng, C.byref(P), 1, stream))      # a failed prefill holds nothing
            try:
                slots = [base + i for i in range(n_copies)]
                status = self._run(eng, slots, n_copies, sp, stream, speculative=True)
                keep = status[0].keep if n_copies > 1 else 0
                rows = self._read_rows(eng, base + keep, status[keep].n_steps, stream)
                if gen_state is not None:
                    gen_state.set_offset(int(status[0].rng_offset))
            finally:
                lib.vcb_release(eng, base, n_copies)
        gen = torch.from_numpy(self._undelay(rows, K)).to(dev)
        res = torch.cat([y[0].long(), gen], dim=1).unsqueeze(0)
        expected = y.shape[2] + rows.shape[0] - K
        assert res.shape == torch.Size((1, K, expected)), f"res.shape: {res.shape}, expected_y_len: {expected}"
        if a.special_first:
            res = res - int(a.n_special)
            gen = gen - int(a.n_special)
        self.last_stats = dict(steps=int(rows.shape[0]), keep=int(keep))
        return res, gen.unsqueeze(0)

    # ------------------------------------------------------------------------------------------------
    # speech editing  (reference voicecraft.py:561-906)
    # ------------------------------------------------------------------------------------------------
    def _edit_prompt(self, y, spans):
        """Segment / placeholder layout of voicecraft.py:239-320, 615-683 as integer index tables.

        y [K,T] (device).  Returns (tokens [T',K] int64 device, mask_rows int32 [T'], more_mask_rows, non_mask)."""
        a = self.args
        K, T = y.shape
        M = len(spans)
        reduced_eog = getattr(a, "reduced_eog", 0)
        starts = [s for s, _ in spans] + [T]
        ends = [0] + [e for _, e in spans]
        non_mask = list(zip(ends, starts))
        # (source interval, end token or None) for every segment, non-masked first then masked
        segs = []
        for i, (s0, s1) in enumerate(non_mask):
            last = i == len(non_mask) - 1
            if a.eos > 0:
                assert reduced_eog
                tail = a.eos if last else None
            elif reduced_eog:
                tail = a.eog if last else None
            else:
                tail = a.eog
            segs.append((s0, s1, tail))
        for (s0, s1) in spans:
            segs.append((s0, s1, a.eog))
        assert not getattr(a, "shuffle_mask_embedding", 0), "shuffle_mask_embedding is a training-time option"
        vals = list(range(a.max_n_spans))[:M]
        mask_val = vals + vals
        # column table: src[k, c] >= 0 -> y[k, src]; otherwise -(token+1)
        cols_src = []
        mask_rows = []
        for j, (s0, s1, tail) in enumerate(segs):
            n_src = (s1 - s0) + (1 if tail is not None else 0)
            blk = np.full((K, n_src + K), -(a.empty_token + 1), dtype=np.int64)
            for k in range(K):
                blk[k, 1 + k: 1 + k + (s1 - s0)] = np.arange(s0, s1)
                if tail is not None:
                    blk[k, 1 + k + (s1 - s0)] = -(tail + 1)
            cols_src.append(blk)
            mask_rows += [-1] * blk.shape[1]
            if j < len(segs) - 1:
                cols_src.append(np.full((K, 1), -(a.eog + 1), dtype=np.int64))    # placeholder column (:264-288)
                mask_rows.append(mask_val[j])
        src = np.concatenate(cols_src, axis=1)
        # cut right after placeholder M plus the first (all-empty) column of the first masked segment (:672-679)
        ph = [i for i, m in enumerate(mask_rows) if m >= 0]
        cut = ph[M] + 2
        src, mask_rows = src[:, :cut], mask_rows[:cut]
        src_t = torch.from_numpy(src).to(y.device)
        tok = torch.where(src_t >= 0, torch.gather(y, 1, src_t.clamp(min=0)), -(src_t + 1))
        return (tok.transpose(1, 0).contiguous(), torch.tensor(mask_rows, dtype=torch.int32, device=y.device),
                mask_val[M + 1:], non_mask)

    @torch.no_grad()
    def inference(self, x: torch.Tensor, x_lens: torch.Tensor, y: torch.Tensor, mask_interval: torch.Tensor,
                  top_k: int = -100, top_p: float = 1.0, temperature: float = 1.0, stop_repetition: int = -1,
                  kvcache: int = 1, silence_tokens: List[int] = [1388, 1898, 131]) -> torch.Tensor:
        a = self.args
        K = a.n_codebooks
        assert x.ndim == 2, x.shape
        assert x_lens.ndim == 1, x_lens.shape
        assert y.ndim == 3, y.shape
        dev = self.mask_embedding.device
        x = x.to(dev)
        y = y.to(dev)
        if a.special_first:
            y = y + int(a.n_special)
        y = y.transpose(2, 1)
        assert y.shape[0] == 1 and y.shape[1] == K, y.shape
        assert mask_interval.shape == torch.Size((1, mask_interval.shape[1], 2)), mask_interval
        spans = [(int(s), int(e)) for s, e in mask_interval[0].tolist()]
        if len(spans) > min(8, int(a.max_n_spans)):
            raise ValueError(f"{len(spans)} masked spans: at most min(8, max_n_spans={a.max_n_spans}) per utterance")
        logging.info(f"silence tokens: {silence_tokens}, note that if you are not using the pretrained encodec "
                     f"6f79c6a8, make sure you specified it yourself, rather than using the default")
        y0 = y[0].long().contiguous()
        y_tok, mask_rows, more_vals, non_mask = self._edit_prompt(y0, spans)
        x_len, y_len = int(x.shape[1]), int(y_tok.shape[0])
        cap = x_len * 10
        need_seq = x_len + max(y_len, cap + 1) + (K + 3) * (len(spans) + 1) + 8
        eng = self._engine(need_slots=1, need_seq=need_seq)
        lib = _lib.load()
        x_ids = x[0].long().contiguous()
        self._check_ids(x_ids, y_tok)
        sp = self._sampling(top_k, top_p, temperature, stop_repetition, silence_tokens)
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream().cuda_stream
            base = self._free_slots(1, self._eng_opts["max_slots"])
            P = _lib.vcb_prompt(slot=base, n_copies=1, mode=1, x_len=x_len, text_ids_dev=x_ids.data_ptr(), y_len=y_len,
                                y_tokens_dev=y_tok.data_ptr(), mask_rows_dev=mask_rows.data_ptr(),
                                n_more_spans=len(more_vals))
            for i, v in enumerate(more_vals):
                P.more_mask_rows[i] = int(v)
            gen_state = self._device_rng(P, dev, 1)
            _lib.check(lib.vcb_prefill(eng, C.byref(P), 1, stream))      # a failed prefill holds nothing
            try:
                status = self._run(eng, [base], 1, sp, stream)
                rows = self._read_rows(eng, base, status[0].n_steps, stream)
                ends = [status[0].span_ends[i] for i in range(status[0].n_spans_done)]
                if gen_state is not None:
                    gen_state.set_offset(int(status[0].rng_offset))
            finally:
                lib.vcb_release(eng, base, 1)
        assert len(ends) == len(spans), f"len(generated): {len(ends)}, num_mask: {len(spans)}"
        pieces, lo = [], 0
        for (s0, s1), hi in zip(non_mask, ends):
            pieces.append(y0[:, s0:s1])
            pieces.append(torch.from_numpy(self._undelay(rows[lo:hi], K)).to(dev))
            lo = hi
        pieces.append(y0[:, non_mask[-1][0]: non_mask[-1][1]])
        res = torch.cat(pieces, dim=1).unsqueeze(0)
        if a.special_first:
            res = res - int(a.n_special)
        self.last_stats = dict(steps=int(rows.shape[0]))
        return res

    # ------------------------------------------------------------------------------------------------
    # driver-level batching of INDEPENDENT utterances (BASELINE config 2).
    # Not a reference API: the reference decodes one utterance per call.  Each utterance keeps its own
    # state machine; one Exp(1) draw of shape [B*K, V] per step feeds all of them.
    # ------------------------------------------------------------------------------------------------
    def open_edit_session(self, xs, ys, mask_intervals, top_k=-100, top_p=1.0, temperature=1.0, stop_repetition=-1,
                          silence_tokens=(1388, 1898, 131), seeds=None, noise_fns=None):
        """Independent speech-editing utterances decoded as one batch (BASELINE config 3).  mask_intervals: list of
        [1,M,2] tensors.  Returns a DecodeSession; results() gives the edited [1,K,T'] per utterance."""
        return DecodeSession(self, xs, ys, self._sampling(top_k, top_p, temperature, stop_repetition, silence_tokens),
                             mask_intervals=mask_intervals, seeds=seeds, noise_fns=noise_fns)

    @torch.no_grad()
    def inference_many(self, xs, ys, mask_intervals, poll_every: int = 4, **kw):
        """Batched counterpart of `inference` (speech editing) for independent utterances."""
        sess = self.open_edit_session(xs, ys, mask_intervals, **kw)
        try:
            sess.sample()
            while True:
                if sess.steps % poll_every == 0 and sess.all_done():
                    break
                sess.step()
            return [r[0] for r in sess.results()]
        finally:
            sess.close()

    def open_tts_session(self, xs, ys, top_k=-100, top_p=1.0, temperature=1.0, stop_repetition=3,
                         silence_tokens=(1388, 1898, 131), seeds=None, noise_fns=None):
        """xs: list of [1,L] int64, ys: list of [1,T,K] int64 (any device).  Prefills every utterance (one packed,
        chunked pass) and returns a DecodeSession whose .step() runs one decode step for all of them.

        seeds: one generator seed per utterance -- utterance i samples from the Philox stream of a torch CUDA generator
        seeded with seeds[i] (offset 0), i.e. its tokens equal ``torch.manual_seed(seeds[i]); inference_tts(x_i, ., y_i)``.
        Default: the device generator's current seed + i at its current offset (the global generator is left untouched).
        noise_fns: instead, one callable(shape=[K,V], device) per utterance (tests: CPU-generator noise for the oracle)."""
        return DecodeSession(self, xs, ys, self._sampling(top_k, top_p, temperature, stop_repetition, silence_tokens),
                             seeds=seeds, noise_fns=noise_fns)

    @torch.no_grad()
    def inference_tts_many(self, xs, ys, poll_every: int = 8, **kw):
        """Returns a list of (res [1,K,T+G], gen [1,K,G]) like inference_tts, one per utterance."""
        sess = self.open_tts_session(xs, ys, **kw)
        try:
            sess.sample()
            while True:
                if sess.steps % poll_every == 0 and sess.all_done():
                    break
                sess.step()
            return sess.results()
        finally:
            sess.close()


class DecodeSession:
    """A batch of independent utterances resident in the engine, one slot and one random stream each."""

    def __init__(self, model: "VoiceCraft", xs, ys, sp, mask_intervals=None, seeds=None, noise_fns=None):
        a = model.args
        self.edit = mask_intervals is not None
        self.non_mask = []
        K = a.n_codebooks
        dev = model.mask_embedding.device
        self.model, self.sp, self.dev, self.K = model, sp, dev, K
        self.B = len(xs)
        self.lib = _lib.load()
        self._open = False
        self.slots = []
        prompts, keep_alive, need_seq = [], [], 0
        self.y0 = []
        if seeds is not None and len(seeds) != self.B:
            raise ValueError("seeds: one per utterance")
        if noise_fns is not None and len(noise_fns) != self.B:
            raise ValueError("noise_fns: one per utterance")
        for x, y in zip(xs, ys):
            assert x.ndim == 2 and y.ndim == 3 and y.shape[2] == K
            x = x.to(dev, non_blocking=True)
            y = y.to(dev, non_blocking=True)
            if a.special_first:
                y = y + int(a.n_special)
            yk = y.transpose(2, 1)[0].long().contiguous()
            idx = len(prompts)
            if self.edit:
                spans = [(int(s), int(e)) for s, e in mask_intervals[idx][0].tolist()]
                if len(spans) > min(8, int(a.max_n_spans)):
                    raise ValueError(f"{len(spans)} masked spans: at most min(8, max_n_spans={a.max_n_spans}) per utterance")
                y_tok, mask_rows, more_vals, non_mask = model._edit_prompt(yk, spans)
                self.non_mask.append(non_mask)
                cap = int(x.shape[1]) * 10
                extra = (K + 3) * (len(spans) + 1)
            else:
                shifted, _ = model.shift([[yk]])
                prompt = shifted[0][0][:, : -(K - 1)] if K > 1 else shifted[0][0]
                y_tok = prompt.transpose(1, 0).contiguous()
                mask_rows, more_vals = None, []
                cap = int(x.shape[1]) * (int(a.encodec_sr) // 5)
                extra = K
            x_ids = x[0].long().contiguous()
            keep_alive += [y_tok, x_ids, mask_rows]
            self.y0.append(yk)
            need_seq = max(need_seq, int(x.shape[1]) + max(int(y_tok.shape[0]), cap + 1) + extra + 8)
            prompts.append((int(x.shape[1]), x_ids, int(y_tok.shape[0]), y_tok, mask_rows, more_vals))
        # one range check for the whole batch (the reference's embedding lookups raise on a bad id)
        model._check_ids(torch.cat([p[1] for p in prompts]), torch.cat([p[3].reshape(-1) for p in prompts]))
        self.eng = model._engine(need_slots=self.B, need_seq=need_seq)
        self.V = model.n_audio_tokens[0]
        base = model._free_slots(self.B, model._eng_opts["max_slots"])
        slots = [base + i for i in range(self.B)]
        # random streams: caller noise (model.noise_fn: one [B*K,V] draw per step; noise_fns: one [K,V] draw per utterance
        # and step) or, by default, one Philox stream per utterance generated inside the sampler kernel
        self._noise_fns = noise_fns
        self._host_noise = noise_fns is not None or model.noise_fn is not None
        self._buf = torch.empty((self.B * K, self.V), device=dev, dtype=torch.float32) if self._host_noise else None
        gen = torch.cuda.default_generators[dev.index or 0]
        seed0, off0 = int(gen.initial_seed()), int(gen.get_offset())
        threads = model._rng_threads(dev, K * self.V)
        P = (_lib.vcb_prompt * self.B)()
        for i, (xl, x_ids, yl, y_tok, mask_rows, more_vals) in enumerate(prompts):
            P[i] = _lib.vcb_prompt(slot=slots[i], n_copies=1, mode=1 if self.edit else 0, x_len=xl, text_ids_dev=x_ids.data_ptr(),
                                   y_len=yl, y_tokens_dev=y_tok.data_ptr(),
                                   mask_rows_dev=mask_rows.data_ptr() if mask_rows is not None else None,
                                   n_more_spans=len(more_vals))
            for j, v in enumerate(more_vals):
                P[i].more_mask_rows[j] = int(v)
            if not self._host_noise:
                P[i].rng_seed = (int(seeds[i]) if seeds is not None else seed0 + i) & 0xFFFFFFFFFFFFFFFF
                P[i].rng_offset = 0 if seeds is not None else off0
                P[i].rng_threads = threads
        self.c_slots = (C.c_int32 * self.B)(*slots)
        self.status = (_lib.vcb_status * self.B)()
        self.steps = 0
        with torch.cuda.device(dev):
            self.stream = torch.cuda.current_stream().cuda_stream
            _lib.check(self.lib.vcb_prefill(self.eng, P, self.B, self.stream))     # a failed prefill holds nothing
        self.slots = slots
        self._open = True
        model._sessions.add(self)
        self._keep_alive = keep_alive

    def _noise(self):
        if not self._host_noise:
            return None
        if self._noise_fns is not None:
            K = self.K
            for i, fn in enumerate(self._noise_fns):
                self._buf[i * K:(i + 1) * K].copy_(fn((K, self.V), self.dev).to(device=self.dev, dtype=torch.float32))
        else:
            self.model._draw_noise(self._buf)
        return self._buf.data_ptr()

    def sample(self):
        """first sampling step (on the prefill's last hidden states)"""
        _lib.check(self.lib.vcb_sample(self.eng, self.c_slots, self.B, self._noise(), C.byref(self.sp), self.stream))
        self.steps += 1

    def step(self):
        # edit sessions: forced hand-over steps of individual utterances ignore their noise rows / consume no draw
        _lib.check(self.lib.vcb_decode_step(self.eng, self.c_slots, self.B, self._noise(), C.byref(self.sp), self.stream))
        self.steps += 1

    def poll(self):
        _lib.check(self.lib.vcb_poll(self.eng, self.c_slots, self.B, self.status, self.stream))
        return self.status

    def all_done(self):
        st = self.poll()
        if any(s.done == 2 for s in st):
            raise _lib.VcbError("decode stopped: engine capacity (max_new_tokens / max_seq_len) exhausted; "
                                "raise it with configure_engine()")
        return all(s.done for s in st)

    def raw_tokens(self, i):
        """delayed token rows [n_steps, K] of utterance i (host numpy)"""
        st = self.poll()
        return self.model._read_rows(self.eng, self.slots[i], st[i].n_steps, self.stream)

    def results(self):
        out = []
        a = self.model.args
        st = self.poll()
        for i in range(self.B):
            rows = self.model._read_rows(self.eng, self.slots[i], st[i].n_steps, self.stream)
            if self.edit:
                assert st[i].done, "edit session results() needs finished utterances"
                ends = [st[i].span_ends[j] for j in range(st[i].n_spans_done)]
                pieces, lo = [], 0
                for (s0, s1), hi in zip(self.non_mask[i], ends):
                    pieces.append(self.y0[i][:, s0:s1])
                    pieces.append(torch.from_numpy(VoiceCraft._undelay(rows[lo:hi], self.K)).to(self.dev))
                    lo = hi
                pieces.append(self.y0[i][:, self.non_mask[i][-1][0]: self.non_mask[i][-1][1]])
                res = torch.cat(pieces, dim=1).unsqueeze(0)
                if a.special_first:
                    res = res - int(a.n_special)
                out.append((res, None))
                continue
            if st[i].done:
                gen = torch.from_numpy(VoiceCraft._undelay(rows, self.K)).to(self.dev)
            else:       # truncated session: drop the still-delayed tail
                n = rows.shape[0]
                gen = torch.from_numpy(np.stack([rows[k: n - (self.K - 1) + k, k] for k in range(self.K)], 0)).to(self.dev) \
                    if n >= self.K else torch.zeros(self.K, 0, dtype=torch.long, device=self.dev)
            res = torch.cat([self.y0[i], gen], dim=1).unsqueeze(0)
            if a.special_first:
                res, gen = res - int(a.n_special), gen - int(a.n_special)
            out.append((res, gen.unsqueeze(0)))
        return out

    def close(self):
        if self._open:
            for s in self.slots:
                self.lib.vcb_release(self.eng, s, 1)
            self._open = False
        self.model._sessions.discard(self)


class ContinuousBatcher:
    """Continuous batching of independent TTS utterances.

    The reference synthesises one utterance per call in a Python loop (inference_tts_scale.py:43-105; the sentence loop of
    gradio_app.py:248-313).  Here up to `max_concurrency` utterances decode together; the moment one finishes, its tokens
    are read, its slot and KV pages are released and the next queued utterance is prefilled into the free slot while the
    others keep decoding.  Every utterance owns its random stream (`seed`), so its result is exactly what
    ``torch.manual_seed(seed); model.inference_tts(x, x_lens, y, ...)`` returns, whatever it was batched with.
    """

    def __init__(self, model: "VoiceCraft", max_concurrency=32, poll_every=8, top_k=-100, top_p=1.0, temperature=1.0,
                 stop_repetition=3, silence_tokens=(1388, 1898, 131)):
        self.model, self.B, self.poll_every = model, int(max_concurrency), max(1, int(poll_every))
        self.sp = model._sampling(top_k, top_p, temperature, stop_repetition, silence_tokens)
        self.queue, self.slots, self._open = [], [], False
        self.stats = dict(steps=0, prefills=0, max_active=0)

    def submit(self, x, y, seed=None):
        """x [1,L] int64, y [1,T,K] int64 (host or device).  Returns the ticket (index into run()'s result list)."""
        self.queue.append((x, y, seed))
        return len(self.queue) - 1

    @torch.no_grad()
    def run(self):
        m, a = self.model, self.model.args
        if m.noise_fn is not None:
            raise _lib.VcbError("ContinuousBatcher uses the per-utterance device generators (model.noise_fn must be None)")
        K, dev, lib = a.n_codebooks, m.mask_embedding.device, _lib.load()
        V = m.n_audio_tokens[0]
        jobs, need_seq = [], 0
        for x, y, seed in self.queue:
            x = x.to(dev, non_blocking=True)
            y = y.to(dev, non_blocking=True)
            if a.special_first:
                y = y + int(a.n_special)
            yk = y.transpose(2, 1)[0].long().contiguous()
            shifted, _ = m.shift([[yk]])
            prompt = shifted[0][0][:, : -(K - 1)] if K > 1 else shifted[0][0]
            y_tok = prompt.transpose(1, 0).contiguous()
            x_ids = x[0].long().contiguous()
            m._check_ids(x_ids, y_tok)
            cap = int(x.shape[1]) * (int(a.encodec_sr) // 5)
            need_seq = max(need_seq, int(x.shape[1]) + max(int(y_tok.shape[0]), cap + 1) + K + 8)
            jobs.append(dict(x_ids=x_ids, y_tok=y_tok, yk=yk, seed=seed))
        n_slots = min(self.B, max(1, len(jobs)))
        eng = m._engine(need_slots=n_slots, need_seq=need_seq)
        base = m._free_slots(n_slots, m._eng_opts["max_slots"])
        self.slots, self._open = [base + i for i in range(n_slots)], True
        m._sessions.add(self)
        gen0 = torch.cuda.default_generators[dev.index or 0]
        seed0, threads = int(gen0.initial_seed()), m._rng_threads(dev, K * V)
        free, active, results, nxt = list(self.slots), {}, [None] * len(jobs), 0
        try:
            with torch.cuda.device(dev):
                stream = torch.cuda.current_stream().cuda_stream
                steps = 0
                while nxt < len(jobs) or active:
                    # ---- refill free slots from the queue: one packed prefill + the first sampling step of the newcomers
                    new = []
                    while free and nxt < len(jobs):
                        new.append((free.pop(0), nxt))
                        nxt += 1
                    if new:
                        P = (_lib.vcb_prompt * len(new))()
                        for j, (slot, ji) in enumerate(new):
                            J = jobs[ji]
                            P[j] = _lib.vcb_prompt(slot=slot, n_copies=1, mode=0, x_len=int(J["x_ids"].shape[0]),
                                                   text_ids_dev=J["x_ids"].data_ptr(), y_len=int(J["y_tok"].shape[0]),
                                                   y_tokens_dev=J["y_tok"].data_ptr(), mask_rows_dev=None, n_more_spans=0)
                            P[j].rng_seed = (int(J["seed"]) if J["seed"] is not None else seed0 + ji) & 0xFFFFFFFFFFFFFFFF
                            P[j].rng_offset = 0
                            P[j].rng_threads = threads
                        _lib.check(lib.vcb_prefill(eng, P, len(new), stream))
                        for slot, ji in new:
                            active[slot] = ji
                        c_new = (C.c_int32 * len(new))(*[s for s, _ in new])
                        _lib.check(lib.vcb_sample(eng, c_new, len(new), None, C.byref(self.sp), stream))
                        self.stats["prefills"] += 1
                    order = sorted(active)
                    c_slots = (C.c_int32 * len(order))(*order)
                    self.stats["max_active"] = max(self.stats["max_active"], len(order))
                    # ---- decode steps for everyone until the next poll
                    for _ in range(self.poll_every):
                        _lib.check(lib.vcb_decode_step(eng, c_slots, len(order), None, C.byref(self.sp), stream))
                        steps += 1
                    status = (_lib.vcb_status * len(order))()
                    _lib.check(lib.vcb_poll(eng, c_slots, len(order), status, stream))
                    for slot, st in zip(order, status):
                        if st.done == 2:
                            raise _lib.VcbError("decode stopped: engine capacity (max_new_tokens / max_seq_len) exhausted; "
                                                "raise it with configure_engine()")
                        if st.done:
                            ji = active.pop(slot)
                            rows = m._read_rows(eng, slot, st.n_steps, stream)
                            gen = torch.from_numpy(VoiceCraft._undelay(rows, K)).to(dev)
                            res = torch.cat([jobs[ji]["yk"], gen], dim=1).unsqueeze(0)
                            if a.special_first:
                                res, gen = res - int(a.n_special), gen - int(a.n_special)
                            results[ji] = (res, gen.unsqueeze(0))
                            lib.vcb_release(eng, slot, 1)
                            free.append(slot)
                self.stats["steps"] = steps
        finally:
            for slot in list(active):
                lib.vcb_release(eng, slot, 1)
            self._open = False
            m._sessions.discard(self)
            self.queue = []
        return results
