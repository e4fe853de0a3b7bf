"""The embedding step in front of the retrieval path, on the same B200.

The reference gets every embedding from an Ollama server over HTTP, one request per text
(`ollama_embed_text`, app/main.py:134-146; `embed_texts_in_batches`, :149-169; `embed_query`,
:172-180; twins in app/embedding_gen.py:143-190).  `GpuEmbeddingEncoder` keeps those three call
signatures and runs the model itself -- mxbai-embed-large is a BERT-large encoder (24 post-LN
layers, hidden 1024, 16 heads, FFN 4096, erf-GELU, CLS pooling, 512 positions, uncased WordPiece) --
so that a query embedding is born in HBM next to the shard it is scored against and ingest no
longer waits on 32,717 HTTP round trips.

Host side (this file, Python like the reference): WordPiece tokeniser, packing of a batch of
sequences into ONE [tokens, 1024] activation matrix (no padding between sequences), the layer
loop.  Device side: five hand-written sm_100a kernels behind the C ABI (`sqe_encoder_*`,
include/sqe_b200.h): tcgen05 GEMMs with fused bias / GELU / residual / QKV-split epilogues, a
one-shot tcgen05 attention kernel, LayerNorm / embedding / pooling row kernels.  There is no CPU
path and no library GEMM.

Weights: a `transformers`-style BertModel state dict (`from_state_dict`, `load` of a torch /
safetensors file) -- or seeded random weights of the same architecture (`random_init`) when, as in
this build environment, the real checkpoint cannot be fetched.  Matmul operands are stored fp16
(the Ollama model file is F16 as well); biases, LayerNorm parameters, the embedding tables, the
residual stream and every LayerNorm are fp32.
"""
from __future__ import annotations

import asyncio
import ctypes
import threading
import unicodedata
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _native as nat

HIDDEN = 1024
HEADS = 16
HEAD_DIM = 64
MAX_TOKENS = nat.SQE_ENC_MAX_TOKENS          # BERT positions
EMBED_DIM = HIDDEN


# ------------------------------------------------------------------------------- tokeniser
def _char_class(ch: str) -> int:
    """0 keep, 1 whitespace, 2 drop (control / NUL / U+FFFD), 3 punctuation, 4 CJK."""
    cp = ord(ch)
    if ch in (" ", "\t", "\n", "\r"):
        return 1
    if cp == 0 or cp == 0xFFFD:
        return 2
    cat = unicodedata.category(ch)
    if cat == "Zs":
        return 1
    if cat[0] == "C":
        return 2
    if 33 <= cp <= 47 or 58 <= cp <= 64 or 91 <= cp <= 96 or 123 <= cp <= 126 or cat[0] == "P":
        return 3
    if (0x4E00 <= cp <= 0x9FFF or 0x3400 <= cp <= 0x4DBF or 0xF900 <= cp <= 0xFAFF or
            0x20000 <= cp <= 0x2A6DF or 0x2A700 <= cp <= 0x2CEAF or 0x2F800 <= cp <= 0x2FA1F):
        return 4
    return 0


class _SplitTable(dict):
    """`str.translate` table filled on demand: code point -> None (drop), ' ' (whitespace),
    ' c ' (punctuation / CJK: a token of its own) or the character itself."""

    def __missing__(self, cp: int):
        ch = chr(cp)
        k = _char_class(ch)
        v = ch if k == 0 else " " if k == 1 else None if k == 2 else " " + ch + " "
        self[cp] = v
        return v


class WordPieceTokenizer:
    """BERT's uncased tokeniser: clean the text, isolate CJK characters and punctuation, lower-case,
    strip accents, then greedy longest-match-first WordPiece (`##` continuation pieces, a word that
    cannot be covered or is longer than 100 characters is one [UNK]).  Words are memoised: natural
    text repeats them, so a chunk costs about one dictionary lookup per word."""

    def __init__(self, vocab: Dict[str, int], *, unk: str = "[UNK]", cls: str = "[CLS]", sep: str = "[SEP]",
                 max_chars_per_word: int = 100, cache_words: int = 1 << 20):
        for t in (unk, cls, sep):
            if t not in vocab:
                raise ValueError(f"vocabulary has no {t} token")
        self.vocab = vocab
        self.unk_id, self.cls_id, self.sep_id = vocab[unk], vocab[cls], vocab[sep]
        self.max_chars = max_chars_per_word
        self._cache: Dict[str, Tuple[int, ...]] = {}
        self._cache_cap = cache_words
        self._table = _SplitTable()

    @classmethod
    def from_file(cls, path: str, **kw) -> "WordPieceTokenizer":
        """A BERT `vocab.txt`: one token per line, the line number is the id."""
        with open(path, encoding="utf-8") as f:
            vocab = {line.rstrip("\n"): i for i, line in enumerate(f)}
        return cls(vocab, **kw)

    @classmethod
    def from_gguf(cls, path_or_file, **kw) -> "WordPieceTokenizer":
        """The vocabulary inside a BERT GGUF file -- the blob an Ollama server holds for
        `mxbai-embed-large` (gguf_model.py restores BERT's own spelling of the pieces)."""
        from . import gguf_model as gm
        if isinstance(path_or_file, gm.GgufFile):
            return cls(gm.wordpiece_vocab(path_or_file), **kw)
        with gm.GgufFile(path_or_file) as g:
            return cls(gm.wordpiece_vocab(g), **kw)

    def _split(self, text: str) -> List[str]:
        """Words and single punctuation / CJK characters, in order.  One `str.translate` (drop
        control characters, blank out whitespace, put spaces around punctuation and CJK) and one
        `split`, both at C speed; the per-character classification runs once per distinct character."""
        return text.translate(self._table).split()

    def _pieces(self, word: str) -> Tuple[int, ...]:
        hit = self._cache.get(word)
        if hit is not None:
            return hit
        # lower-case + strip accents; the result may contain NEW punctuation-free pieces only, but
        # decomposition can expose punctuation (rare): split again in that case
        norm = "".join(c for c in unicodedata.normalize("NFD", word) if unicodedata.category(c) != "Mn").lower()
        out: List[int] = []
        for sub in (self._split(norm) if norm != word else [norm]):
            out.extend(self._wordpiece(sub))
        res = tuple(out)
        if len(self._cache) < self._cache_cap:
            self._cache[word] = res
        return res

    def _wordpiece(self, word: str) -> List[int]:
        if not word:
            return []
        if len(word) > self.max_chars:
            return [self.unk_id]
        vocab = self.vocab
        pieces: List[int] = []
        start, n = 0, len(word)
        while start < n:
            end = n
            found = None
            while start < end:
                sub = word[start:end] if start == 0 else "##" + word[start:end]
                found = vocab.get(sub)
                if found is not None:
                    break
                end -= 1
            if found is None:
                return [self.unk_id]
            pieces.append(found)
            start = end
        return pieces

    def tokenize(self, text: str) -> List[int]:
        """Piece ids without [CLS] / [SEP]."""
        ids: List[int] = []
        for w in self._split(text):
            ids.extend(self._pieces(w))
        return ids

    def encode(self, text: str, max_tokens: int = MAX_TOKENS) -> List[int]:
        """[CLS] pieces [SEP], truncated to the model's positions (Ollama truncates to the context too)."""
        return [self.cls_id] + self.tokenize(text)[: max_tokens - 2] + [self.sep_id]


# --------------------------------------------------------------------------------- weights
class EncoderWeights:
    """Device-resident parameters.  Per layer: Wqkv [3072,1024] fp16 (query | key | value rows of the
    three Linear weights), Wo [1024,1024], W1 [4096,1024], W2 [1024,4096] fp16; biases and LayerNorm
    parameters fp32."""

    def __init__(self, device: torch.device):
        self.device = device
        self.layers: List[Dict[str, torch.Tensor]] = []
        self.word = self.position = self.type0 = self.emb_g = self.emb_b = None
        self.vocab_size = 0
        self.max_pos = 0
        self.eps = 1e-12

    @classmethod
    def from_state_dict(cls, sd: Dict[str, torch.Tensor], device=None, eps: float = 1e-12) -> "EncoderWeights":
        """`sd`: BertModel names, with or without a `bert.` prefix."""
        dev = torch.device(device if device is not None else "cuda")
        if dev.type != "cuda":
            raise RuntimeError("sqe_b200 has no CPU path: the encoder needs a CUDA device")
        if any(k.startswith("bert.") for k in sd):
            sd = {k[5:]: v for k, v in sd.items() if k.startswith("bert.")}
        w = cls(dev)
        w.eps = float(eps)

        def f32(name):
            return sd[name].detach().to(device=dev, dtype=torch.float32).contiguous()

        def f16(t):
            return t.detach().to(device=dev, dtype=torch.float16).contiguous()

        w.word = f32("embeddings.word_embeddings.weight")
        w.position = f32("embeddings.position_embeddings.weight")
        w.type0 = f32("embeddings.token_type_embeddings.weight")[0].contiguous()
        w.emb_g = f32("embeddings.LayerNorm.weight")
        w.emb_b = f32("embeddings.LayerNorm.bias")
        w.vocab_size, hidden = w.word.shape
        w.max_pos = w.position.shape[0]
        if hidden != HIDDEN:
            raise ValueError(f"hidden size {hidden}: the kernels are built for {HIDDEN} (mxbai-embed-large)")
        i = 0
        while f"encoder.layer.{i}.attention.self.query.weight" in sd:
            p = f"encoder.layer.{i}."
            wq, wk, wv = (sd[p + f"attention.self.{n}.weight"] for n in ("query", "key", "value"))
            bq, bk, bv = (sd[p + f"attention.self.{n}.bias"] for n in ("query", "key", "value"))
            layer = {
                "wqkv": f16(torch.cat([wq, wk, wv], dim=0)),
                "bqkv": torch.cat([bq, bk, bv]).detach().to(device=dev, dtype=torch.float32).contiguous(),
                "wo": f16(sd[p + "attention.output.dense.weight"]), "bo": f32(p + "attention.output.dense.bias"),
                "g1": f32(p + "attention.output.LayerNorm.weight"), "b1": f32(p + "attention.output.LayerNorm.bias"),
                "w1": f16(sd[p + "intermediate.dense.weight"]), "bi": f32(p + "intermediate.dense.bias"),
                "w2": f16(sd[p + "output.dense.weight"]), "bo2": f32(p + "output.dense.bias"),
                "g2": f32(p + "output.LayerNorm.weight"), "b2": f32(p + "output.LayerNorm.bias"),
            }
            if layer["wqkv"].shape != (3 * HIDDEN, HIDDEN) or layer["w1"].shape[1] != HIDDEN or \
                    layer["w1"].shape[0] % 256 != 0 or layer["w2"].shape != (HIDDEN, layer["w1"].shape[0]):
                raise ValueError("unsupported layer geometry")
            w.layers.append(layer)
            i += 1
        if not w.layers:
            raise ValueError("state dict has no encoder layers")
        return w

    @classmethod
    def load(cls, path: str, device=None) -> "EncoderWeights":
        """A `pytorch_model.bin` / `.pt` state dict or a `.safetensors` file of a BERT encoder."""
        if path.endswith(".safetensors"):
            from safetensors.torch import load_file
            sd = load_file(path)
        else:
            sd = torch.load(path, map_location="cpu", weights_only=True)
        return cls.from_state_dict(sd, device=device)

    @classmethod
    def from_gguf(cls, path_or_file, device=None) -> "EncoderWeights":
        """A BERT GGUF file (what `ollama pull mxbai-embed-large` stores; F32 / F16 / BF16 / Q8_0 / Q4_0 /
        Q4_1 tensors).  The file's hyper-parameters must be the ones the kernels are built for: 16 heads of
        64, hidden 1024, CLS pooling, bidirectional attention."""
        from . import gguf_model as gm
        g = path_or_file if isinstance(path_or_file, gm.GgufFile) else gm.GgufFile(path_or_file)
        try:
            sd, cfg = gm.bert_state_dict(g)
        finally:
            if g is not path_or_file:
                g.close()
        if cfg["heads"] != HEADS or cfg["hidden"] != HIDDEN:
            raise ValueError(f"{cfg['heads']} heads x hidden {cfg['hidden']}: the kernels are built for "
                             f"{HEADS} x {HIDDEN} (mxbai-embed-large)")
        if cfg["pooling"] not in (gm.POOLING_CLS, gm.POOLING_NONE):
            raise ValueError(f"pooling type {cfg['pooling']}: the encoder returns the [CLS] state (pooling type "
                             f"{gm.POOLING_CLS}, what mxbai-embed-large uses)")
        if cfg["causal"]:
            raise ValueError("causal attention: not a BERT encoder")
        if cfg["max_positions"] > MAX_TOKENS:
            raise ValueError(f"{cfg['max_positions']} positions: the kernels are built for {MAX_TOKENS}")
        return cls.from_state_dict(sd, device=device, eps=cfg["eps"])

    @classmethod
    def random_init(cls, seed: int = 0, layers: int = 24, device=None, vocab: int = 30522,
                    intermediate: int = 4096) -> "EncoderWeights":
        """Random weights of the mxbai-embed-large architecture, generated on the device (the real
        checkpoint cannot be fetched in an offline environment; bench.py says so in `data`)."""
        dev = torch.device(device if device is not None else "cuda")
        g = torch.Generator(device=dev).manual_seed(seed)

        def rnd(*shape, s=0.03):
            return torch.randn(*shape, generator=g, device=dev, dtype=torch.float32) * s

        sd = {
            "embeddings.word_embeddings.weight": rnd(vocab, HIDDEN, s=0.5),
            "embeddings.position_embeddings.weight": rnd(MAX_TOKENS, HIDDEN, s=0.3),
            "embeddings.token_type_embeddings.weight": rnd(2, HIDDEN, s=0.1),
            "embeddings.LayerNorm.weight": 1.0 + rnd(HIDDEN, s=0.1),
            "embeddings.LayerNorm.bias": rnd(HIDDEN, s=0.1),
        }
        for i in range(layers):
            p = f"encoder.layer.{i}."
            for n in ("query", "key", "value"):
                sd[p + f"attention.self.{n}.weight"] = rnd(HIDDEN, HIDDEN, s=0.05 if n != "value" else 0.03)
                sd[p + f"attention.self.{n}.bias"] = rnd(HIDDEN, s=0.1)
            sd[p + "attention.output.dense.weight"] = rnd(HIDDEN, HIDDEN)
            sd[p + "attention.output.dense.bias"] = rnd(HIDDEN, s=0.1)
            sd[p + "attention.output.LayerNorm.weight"] = 1.0 + rnd(HIDDEN, s=0.1)
            sd[p + "attention.output.LayerNorm.bias"] = rnd(HIDDEN, s=0.1)
            sd[p + "intermediate.dense.weight"] = rnd(intermediate, HIDDEN)
            sd[p + "intermediate.dense.bias"] = rnd(intermediate, s=0.1)
            sd[p + "output.dense.weight"] = rnd(HIDDEN, intermediate)
            sd[p + "output.dense.bias"] = rnd(HIDDEN, s=0.1)
            sd[p + "output.LayerNorm.weight"] = 1.0 + rnd(HIDDEN, s=0.1)
            sd[p + "output.LayerNorm.bias"] = rnd(HIDDEN, s=0.1)
        return cls.from_state_dict(sd, device=dev)

    @property
    def intermediate(self) -> int:
        return self.layers[0]["w1"].shape[0]

    def c_struct(self) -> "nat.SqeEncoderWeights":
        """The `SqeEncoderWeights` view of these tensors for `sqe_encoder_forward` (built once)."""
        if getattr(self, "_c", None) is None:
            arr = (nat.SqeEncoderLayer * len(self.layers))()
            names = (("wqkv", "wqkv"), ("bqkv", "bqkv"), ("wo", "wo"), ("bo", "bo"), ("ln1_gamma", "g1"), ("ln1_beta", "b1"),
                     ("w1", "w1"), ("b1", "bi"), ("w2", "w2"), ("b2", "bo2"), ("ln2_gamma", "g2"), ("ln2_beta", "b2"))
            for i, L in enumerate(self.layers):
                for field, key in names:
                    setattr(arr[i], field, L[key].data_ptr())
            c = nat.SqeEncoderWeights(len(self.layers), self.vocab_size, self.max_pos, self.intermediate, self.eps,
                                      self.word.data_ptr(), self.position.data_ptr(), self.type0.data_ptr(),
                                      self.emb_g.data_ptr(), self.emb_b.data_ptr(), arr)
            self._c = (c, arr)                              # keep the layer array alive
        return self._c[0]

    def flops_per_token(self) -> float:
        """Linear layers only (2 FLOP per multiply-add); attention adds 4 * S * 1024 per token."""
        return 2.0 * len(self.layers) * (4 * HIDDEN * HIDDEN + 2 * HIDDEN * self.intermediate)


# ------------------------------------------------------------------------------ tensor ops
def _ptr(t: Optional[torch.Tensor]) -> int:
    return 0 if t is None else t.data_ptr()


def gemm(x: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, epilogue: int, out0: torch.Tensor, *,
         m: Optional[int] = None, out1: Optional[torch.Tensor] = None, n_split: int = 0, q_cols: int = 0,
         q_scale: float = 1.0, residual: Optional[torch.Tensor] = None, res_stats: Optional[torch.Tensor] = None,
         res_gamma: Optional[torch.Tensor] = None, res_beta: Optional[torch.Tensor] = None) -> None:
    """`sqe_encoder_gemm` on the current stream: out0 (+ out1) = epilogue(x[:m] @ w.T + bias).  With
    `res_stats` ([rows, 2] = mean, rstd) the residual operand is LayerNorm(residual) recomputed on the fly."""
    dev = x.device
    if not x.is_cuda:
        raise RuntimeError("sqe_b200 has no CPU path: tensors must live on a CUDA device")
    m = x.shape[0] if m is None else m
    n, k = w.shape
    with torch.cuda.device(dev):
        nat.call("sqe_encoder_gemm", x.data_ptr(), x.stride(0), w.data_ptr(), bias.data_ptr(), m, n, k, epilogue,
                 out0.data_ptr(), out0.stride(0), _ptr(out1), 0 if out1 is None else out1.stride(0), n_split, q_cols,
                 q_scale, _ptr(residual), 0 if residual is None else residual.stride(0), _ptr(res_stats), _ptr(res_gamma),
                 _ptr(res_beta), torch.cuda.current_stream(dev).cuda_stream)


def layernorm(x: torch.Tensor, g: torch.Tensor, b: torch.Tensor, eps: float, out32: Optional[torch.Tensor],
              out16: torch.Tensor, rows: Optional[int] = None, stats: Optional[torch.Tensor] = None) -> None:
    """`sqe_encoder_layernorm`; with `stats` ([rows, 2]) the row statistics are written and out32 may be None."""
    dev = x.device
    with torch.cuda.device(dev):
        nat.call("sqe_encoder_layernorm", x.data_ptr(), g.data_ptr(), b.data_ptr(), eps,
                 x.shape[0] if rows is None else rows, _ptr(out32), out16.data_ptr(), _ptr(stats),
                 torch.cuda.current_stream(dev).cuda_stream)


def attention(qk: torch.Tensor, vt: torch.Tensor, tiles: torch.Tensor, n_tiles: int, max_len: int,
              ctx: torch.Tensor) -> None:
    dev = qk.device
    with torch.cuda.device(dev):
        nat.call("sqe_encoder_attention", qk.data_ptr(), vt.data_ptr(), qk.shape[0], tiles.data_ptr(), n_tiles,
                 max_len, ctx.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)


class _Buffers:
    """Activation buffers of one token capacity (a multiple of 128 rows; V^T's row pitch depends on it)."""

    def __init__(self, t_pad: int, inter: int, dev: torch.device):
        z = lambda *s, dt: torch.zeros(*s, dtype=dt, device=dev)          # noqa: E731
        self.t_pad = t_pad
        self.sum_a = z(t_pad, HIDDEN, dt=torch.float32)    # pre-LayerNorm sums = the fp32 residual stream (ping-pong):
        self.sum_b = z(t_pad, HIDDEN, dt=torch.float32)    # a LayerNorm's fp32 output is recomputed where it is needed
        self.stats_a = z(t_pad, 2, dt=torch.float32)       # {mean, rstd} of the rows of sum_a / sum_b
        self.stats_b = z(t_pad, 2, dt=torch.float32)
        self.h16 = z(t_pad, HIDDEN, dt=torch.float16)      # LayerNorm output as the next tensor-core operand
        self.qk = z(t_pad, 2 * HIDDEN, dt=torch.float16)   # Q / 8 | K
        self.vt = z(HIDDEN, t_pad, dt=torch.float16)       # V^T
        self.ctx = z(t_pad, HIDDEN, dt=torch.float16)
        self.ffn = z(t_pad, inter, dt=torch.float16)
        # per-forward metadata, one H2D copy: ids | pos | first_token | tiles
        self.meta_cap = 2 * t_pad + t_pad + 4 * (t_pad // 128 + t_pad)
        self.meta_dev = torch.zeros(self.meta_cap, dtype=torch.int32, device=dev)
        # two pinned staging buffers, alternating: the copy of forward i may still be queued when
        # the host fills the metadata of forward i + 1
        self.meta_host = [torch.zeros(self.meta_cap, dtype=torch.int32).pin_memory() for _ in range(2)]
        self.meta_done = [None, None]
        self.turn = 0
        # query-sized capacities carry the workspace of the few-token GEMM (partial tiles + tickets)
        self.small_ws = None
        if t_pad <= 128:
            self.small_ws = torch.zeros(int(nat.load().sqe_encoder_gemm_small_workspace_bytes()), dtype=torch.uint8, device=dev)
        self.c = nat.SqeEncoderBuffers(t_pad, self.sum_a.data_ptr(), self.sum_b.data_ptr(), self.stats_a.data_ptr(),
                                       self.stats_b.data_ptr(), self.h16.data_ptr(),
                                       self.qk.data_ptr(), self.vt.data_ptr(), self.ctx.data_ptr(), self.ffn.data_ptr(),
                                       0 if self.small_ws is None else self.small_ws.data_ptr(),
                                       0 if self.small_ws is None else self.small_ws.numel())
        self.graph_out = torch.zeros(t_pad // 8, HIDDEN, dtype=torch.float32, device=dev)     # one row per sequence
        self.graphs: Dict[tuple, "torch.cuda.CUDAGraph"] = {}
        self.seen: Dict[tuple, int] = {}


class GpuEmbeddingEncoder:
    """`embed_texts(texts) -> np.ndarray [n, 1024]` and the reference's three coroutines.

    `normalize=False` returns the raw CLS hidden state (what an embedding server returns is
    normalised downstream anyway: app/main.py:315-316, :353-354 divide by the norm again)."""

    def __init__(self, weights: EncoderWeights, tokenizer: Optional[WordPieceTokenizer] = None, *,
                 max_batch_tokens: int = 32768, blank_policy: str = "main", use_graphs: bool = True,
                 graph_max_tokens: int = 256):
        self.w = weights
        self.device = weights.device
        self.tok = tokenizer
        self.max_batch_tokens = max(int(max_batch_tokens), MAX_TOKENS)
        self.blank_policy = blank_policy
        self._bufs: Dict[Tuple[int, int], _Buffers] = {}
        self._lock = threading.Lock()
        self.stream = torch.cuda.Stream(device=self.device)
        self.launches_last_forward = 0
        # a query is a handful of tokens: its 170 launches are captured in a CUDA graph the second
        # time a batch shape (token capacity, entries, longest sequence, sequences) shows up
        self.use_graphs = use_graphs
        self.graph_max_tokens = graph_max_tokens
        self.graph_replays = 0
        nat.load()                                          # fail loudly when the CUDA library is missing

    @classmethod
    def from_gguf(cls, path: str, device=None, **kw) -> "GpuEmbeddingEncoder":
        """Weights AND vocabulary from one BERT GGUF file."""
        from . import gguf_model as gm
        with gm.GgufFile(path) as g:
            tok = WordPieceTokenizer.from_gguf(g)
            w = EncoderWeights.from_gguf(g, device=device)
        return cls(w, tok, **kw)

    @classmethod
    def from_ollama(cls, name: str = "mxbai-embed-large", models_dir: Optional[str] = None, device=None,
                    **kw) -> "GpuEmbeddingEncoder":
        """The model the reference asks its Ollama server for (`EMBED_MODEL_NAME`, app/main.py:29),
        loaded from that server's model store (`~/.ollama/models` or $OLLAMA_MODELS) instead of
        being called over HTTP."""
        from . import gguf_model as gm
        return cls.from_gguf(gm.find_ollama_model(name, models_dir), device=device, **kw)

    # ------------------------------------------------------------------ device forward
    def _buffers(self, t_pad: int) -> _Buffers:
        """Activation buffers for this token capacity ON THE CURRENT STREAM: a forward pass only
        orders itself against earlier work of its own stream, so two streams (say the ingest thread's
        `embed_texts` and a MicroBatcher's compute stream) must never share a set."""
        key = (t_pad, torch.cuda.current_stream(self.device).cuda_stream)
        b = self._bufs.get(key)
        if b is None:
            if len(self._bufs) >= 6:                        # keep the few capacities that recur
                old = next(iter(self._bufs))
                if old[1] != key[1]:                        # another stream's set may still be in use there
                    torch.cuda.synchronize(self.device)
                self._bufs.pop(old)
            b = self._bufs[key] = _Buffers(t_pad, self.w.intermediate, self.device)
        return b

    @staticmethod
    def plan(lengths: Sequence[int]) -> Tuple[int, np.ndarray, np.ndarray, np.ndarray]:
        """Packing of sequences of the given lengths: (t_pad, pos [t_pad], first_token [n], entries [m,4]).
        Every sequence starts at a multiple of 8 tokens: the attention kernel fetches V^T with TMA
        boxes whose innermost coordinate is the token index, and a box must start on a 16-byte
        boundary.  pos = -1 marks the (zero) filler rows."""
        lens = np.asarray(lengths, dtype=np.int64)
        if lens.size == 0 or lens.min() < 1 or lens.max() > MAX_TOKENS:
            raise ValueError(f"sequence lengths must be in [1, {MAX_TOKENS}]")
        slots = (lens + 7) // 8 * 8
        starts = np.concatenate([[0], np.cumsum(slots)[:-1]])
        t_pad = (int(slots.sum()) + 127) // 128 * 128
        pos = np.full(t_pad, -1, dtype=np.int32)
        rows = np.repeat(starts, lens) + (np.arange(int(lens.sum())) - np.repeat(np.cumsum(lens) - lens, lens))
        pos[rows] = rows - np.repeat(starts, lens)
        # attention work list: (first token, length, first query row, query rows).  One CTA keeps the
        # K / V of its (sequence, head) in shared memory and walks `group` 128-query tiles; small
        # batches get one tile per entry (more CTAs), ingest batches whole sequences (K / V fetched once).
        n_q = (lens + 127) // 128
        group = int(min(4, max(1, (int(n_q.sum()) * HEADS) // (3 * 148))))
        n_e = (n_q + group - 1) // group
        seq_of = np.repeat(np.arange(lens.size), n_e)
        first_e = np.concatenate([[0], np.cumsum(n_e)[:-1]])
        q0 = (np.arange(seq_of.size) - first_e[seq_of]) * (128 * group)
        tiles = np.zeros((seq_of.size, 4), dtype=np.int32)
        tiles[:, 0] = starts[seq_of]
        tiles[:, 1] = lens[seq_of]
        tiles[:, 2] = q0
        tiles[:, 3] = np.minimum(128 * group, lens[seq_of] - q0)
        return t_pad, pos, starts.astype(np.int32), tiles

    def forward_ids(self, seqs: Sequence[Sequence[int]], out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """One packed forward pass on the current stream: CLS hidden states [n, 1024] fp32 on the
        device.  The caller bounds the batch (`embed_token_ids` splits by `max_batch_tokens`)."""
        n = len(seqs)
        lens = [len(s) for s in seqs]
        t_pad, pos, first, tiles = self.plan(lens)
        total = int(sum(lens))
        dev = self.device
        rows = np.flatnonzero(pos >= 0)                     # token rows (the rest is filler: id -1 -> zeros)
        with self._lock:
            b = self._buffers(t_pad)
            n_tiles = tiles.shape[0]
            slot = b.turn
            b.turn ^= 1
            if b.meta_done[slot] is not None:
                b.meta_done[slot].synchronize()             # its previous copy has left the buffer
            mh = b.meta_host[slot].numpy()
            o_pos, o_first, o_tiles = t_pad, 2 * t_pad, 2 * t_pad + ((n + 3) // 4) * 4
            used = o_tiles + 4 * n_tiles
            if used > b.meta_cap:
                raise ValueError("too many sequences for this token capacity")
            mh[:t_pad] = -1
            mh[rows] = np.fromiter((t for s in seqs for t in s), dtype=np.int32, count=total)
            mh[o_pos:o_pos + t_pad] = np.maximum(pos, 0)
            mh[o_first:o_first + n] = first
            mh[o_tiles:used] = tiles.reshape(-1)
            b.meta_dev[:used].copy_(b.meta_host[slot][:used], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(dev))
            b.meta_done[slot] = ev
            ids_d, pos_d = b.meta_dev[:t_pad], b.meta_dev[o_pos:o_pos + t_pad]
            first_d, tiles_d = b.meta_dev[o_first:o_first + n], b.meta_dev[o_tiles:used]
            if out is None:
                out = torch.empty((n, HIDDEN), dtype=torch.float32, device=dev)
            max_len = max(lens)
            rows_used = int(first[-1]) + lens[-1]           # token rows in use (the last sequence's end)
            key = (n_tiles, (max_len + 63) // 64, n, o_first, o_tiles, (rows_used + 15) // 16)
            graph = None
            if self.use_graphs and t_pad <= self.graph_max_tokens:
                graph = b.graphs.get(key)
                b.seen[key] = b.seen.get(key, 0) + 1
                if graph is None and b.seen[key] == 2 and len(b.graphs) < 16:
                    graph = self._capture(b, ids_d, pos_d, tiles_d, n_tiles, max_len, first_d, n, key, rows_used)
            if graph is not None:
                graph.replay()
                self.graph_replays += 1
                out.copy_(b.graph_out[:n])
                self.launches_last_forward = 2 + 7 * len(self.w.layers)
            else:
                self._forward(b, ids_d, pos_d, tiles_d, n_tiles, max_len, first_d, n, out, rows_used)
        return out

    def _forward(self, b: _Buffers, ids_d, pos_d, tiles_d, n_tiles: int, max_len: int, first_d, n: int,
                 out: torch.Tensor, rows_used: int = 0) -> None:
        """`sqe_encoder_forward` on the current stream: 2 + 7 layers kernel launches, one call."""
        dev = self.device
        with torch.cuda.device(dev):
            nat.call("sqe_encoder_forward", ctypes.addressof(self.w.c_struct()), ctypes.addressof(b.c),
                     ids_d.data_ptr(), pos_d.data_ptr(), tiles_d.data_ptr(), n_tiles, max_len, first_d.data_ptr(), n,
                     rows_used, out.data_ptr(), out.stride(0), torch.cuda.current_stream(dev).cuda_stream)
        self.launches_last_forward = 2 + 7 * len(self.w.layers)
        nat.launch_count += self.launches_last_forward

    def _capture(self, b: _Buffers, ids_d, pos_d, tiles_d, n_tiles: int, max_len: int, first_d, n: int, key,
                 rows_used: int = 0):
        """Capture the forward pass of this batch shape (it reads the metadata from fixed device
        addresses, so a replay serves any batch of the same shape).  None if capture fails."""
        dev = self.device
        cur = torch.cuda.current_stream(dev)
        try:
            self._forward(b, ids_d, pos_d, tiles_d, n_tiles, max_len, first_d, n, b.graph_out, rows_used)     # warm: attributes set
            cur.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=self._capture_stream(), capture_error_mode="thread_local"):
                self._forward(b, ids_d, pos_d, tiles_d, n_tiles, max_len, first_d, n, b.graph_out, rows_used)
        except Exception as e:                              # noqa: BLE001 -- the eager path still works
            print(f"[sqe_b200] encoder graph capture failed ({e}); staying on direct launches")
            self.use_graphs = False
            return None
        b.graphs[key] = g
        return g

    def _capture_stream(self) -> "torch.cuda.Stream":
        if getattr(self, "_cap_stream", None) is None:
            self._cap_stream = torch.cuda.Stream(device=self.device)
        return self._cap_stream

    # ----------------------------------------------------------------------- batching
    def _batches(self, lens: Sequence[int]) -> Iterable[Tuple[int, int]]:
        i, n = 0, len(lens)
        while i < n:
            j, tok = i, 0
            while j < n and tok + lens[j] <= self.max_batch_tokens:
                tok += lens[j]
                j += 1
            yield i, j
            i = j

    def embed_token_ids(self, seqs: Sequence[Sequence[int]]) -> torch.Tensor:
        """[n, 1024] fp32 on the device (asynchronous on `self.stream`; synchronise before reading)."""
        n = len(seqs)
        out = torch.empty((n, HIDDEN), dtype=torch.float32, device=self.device)
        if n == 0:
            return out
        lens = [len(s) for s in seqs]
        out.record_stream(self.stream)                      # allocated on the caller's stream, written on ours
        with torch.cuda.stream(self.stream):
            for i, j in self._batches(lens):
                self.forward_ids(seqs[i:j], out=out[i:j])
        return out

    def embed_texts_device(self, texts: Sequence[str]) -> torch.Tensor:
        """[n, 1024] fp32 on the device (asynchronous on `self.stream`).  Texts are tokenised batch by
        batch: while the GPU runs the forward pass of one batch of up to `max_batch_tokens` tokens the
        host tokenises and packs the next, so an ingest runs at the slower of the two, not their sum."""
        if self.tok is None:
            raise RuntimeError("this encoder was built without a tokenizer (pass token ids instead)")
        n = len(texts)
        out = torch.empty((n, HIDDEN), dtype=torch.float32, device=self.device)
        if n == 0:
            return out
        out.record_stream(self.stream)
        with torch.cuda.stream(self.stream):
            first, batch, tokens = 0, [], 0
            for i, text in enumerate(texts):
                ids = self.tok.encode(text)
                if batch and tokens + len(ids) > self.max_batch_tokens:
                    self.forward_ids(batch, out=out[first:i])
                    first, batch, tokens = i, [], 0
                batch.append(ids)
                tokens += len(ids)
            self.forward_ids(batch, out=out[first:n])
        return out

    def embed_texts(self, texts: Sequence[str]) -> np.ndarray:
        """Host array [n, 1024] fp32; blank texts give zero rows (embedding_gen.py:147-148)."""
        texts = list(texts)
        out = np.zeros((len(texts), HIDDEN), dtype=np.float32)
        keep = [i for i, t in enumerate(texts) if t.strip()]
        if keep:
            dev = self.embed_texts_device([texts[i] for i in keep])
            self.stream.synchronize()
            out[keep] = dev.cpu().numpy()
        return out

    # ------------------------------------------------- the reference's call signatures
    async def ollama_embed_text(self, text: str, model: str = "") -> List[float]:
        """app/main.py:134-146 / app/embedding_gen.py:143-166 (`model` is accepted and ignored: this
        object IS the model).  Blank text: a zero vector, as the upload service returns."""
        return self.embed_texts([text])[0].tolist()

    async def embed_texts_in_batches(self, texts: List[str], batch_size: int = 64) -> np.ndarray:
        """app/main.py:149-169: `np.array([])` for no texts; app/embedding_gen.py:169-190 returns
        `zeros((0, 1024))` -- `blank_policy` picks which module's convention applies."""
        if not texts:
            return np.array([]) if self.blank_policy == "main" else np.zeros((0, HIDDEN), dtype=np.float32)
        loop = asyncio.get_running_loop()
        return await loop.run_in_executor(None, self.embed_texts, texts)

    async def embed_query(self, query: str) -> np.ndarray:
        """app/main.py:172-180: `np.array([])` for a blank query, else shape (1, 1024) fp32."""
        if not query.strip():
            return np.array([])
        return self.embed_texts([query])


def install_encoder(module, encoder: GpuEmbeddingEncoder) -> GpuEmbeddingEncoder:
    """Patch a loaded reference module (`main` or `embedding_gen`) so that its embedding calls run
    on the GPU instead of going to the Ollama server."""
    encoder.blank_policy = "embedding_gen" if hasattr(module, "bulk_index_embeddings") else "main"
    module.ollama_embed_text = encoder.ollama_embed_text
    module.embed_texts_in_batches = encoder.embed_texts_in_batches
    if hasattr(module, "embed_query"):
        module.embed_query = encoder.embed_query
    module._sqe_b200_encoder = encoder
    return encoder
