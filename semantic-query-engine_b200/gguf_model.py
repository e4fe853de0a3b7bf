"""The Ollama deployment's own model file as the encoder checkpoint.

The reference never holds the embedding model: it POSTs every text to an Ollama server
(`ollama_embed_text`, app/main.py:134-146, model name `EMBED_MODEL_NAME` = "mxbai-embed-large",
app/main.py:29) and Ollama serves the model from a GGUF blob in its model store.  A deployment that
switches to `GpuEmbeddingEncoder` (encoder.py) therefore already HAS the weights and the vocabulary on
disk -- in GGUF, not as a `transformers` checkpoint.  This module reads that file:

  * `GgufFile`            the container (GGUF v2 / v3, little endian): key / value metadata, tensor
                          directory, tensor data as numpy arrays (F32 / F16 / BF16 as stored, Q8_0 /
                          Q4_0 / Q4_1 blocks dequantised to fp32)
  * `bert_state_dict`     the tensors of a `general.architecture = "bert"` file under the BertModel
                          names `EncoderWeights.from_state_dict` takes (llama.cpp's converter renames
                          them; the table below is the inverse of its BERT tensor map)
  * `wordpiece_vocab`     `tokenizer.ggml.tokens` back to a BERT `vocab.txt` mapping (the converter
                          marks word-initial pieces with U+2581 and strips the `##` of continuation
                          pieces; specials in brackets are kept)
  * `find_ollama_model`   "mxbai-embed-large[:tag]" -> the blob path, through the manifest of Ollama's
                          model store (`~/.ollama/models`, or $OLLAMA_MODELS)

Host-side parsing only (the file format of a third party, restated from its public specification and
checked in tests/test_gguf_model.py against files written by the `gguf` library's own writer); nothing
here computes embeddings and nothing here is a CPU path of the encoder.
"""
from __future__ import annotations

import json
import mmap
import os
import struct
from typing import Any, Dict, List, Optional, Tuple

import numpy as np

GGUF_MAGIC = 0x46554747                      # "GGUF", little endian
_DEFAULT_ALIGNMENT = 32

# metadata value types
_T_U8, _T_I8, _T_U16, _T_I16, _T_U32, _T_I32, _T_F32, _T_BOOL, _T_STR, _T_ARR, _T_U64, _T_I64, _T_F64 = range(13)
_SCALAR = {_T_U8: "<B", _T_I8: "<b", _T_U16: "<H", _T_I16: "<h", _T_U32: "<I", _T_I32: "<i", _T_F32: "<f",
           _T_BOOL: "<?", _T_U64: "<Q", _T_I64: "<q", _T_F64: "<d"}
_NP_SCALAR = {_T_U8: "u1", _T_I8: "i1", _T_U16: "<u2", _T_I16: "<i2", _T_U32: "<u4", _T_I32: "<i4", _T_F32: "<f4",
              _T_BOOL: "?", _T_U64: "<u8", _T_I64: "<i8", _T_F64: "<f8"}

# ggml tensor types this reader understands: id -> (name, elements per block, bytes per block)
GGML_TYPES = {0: ("F32", 1, 4), 1: ("F16", 1, 2), 30: ("BF16", 1, 2),
              8: ("Q8_0", 32, 34), 2: ("Q4_0", 32, 18), 3: ("Q4_1", 32, 20)}

POOLING_NONE, POOLING_MEAN, POOLING_CLS, POOLING_LAST = 0, 1, 2, 3


class GgufError(ValueError):
    pass


class GgufTensorInfo:
    __slots__ = ("name", "shape", "ggml_type", "offset", "nbytes")

    def __init__(self, name: str, shape: Tuple[int, ...], ggml_type: int, offset: int, nbytes: int):
        self.name, self.shape, self.ggml_type, self.offset, self.nbytes = name, shape, ggml_type, offset, nbytes

    def __repr__(self) -> str:
        t = GGML_TYPES.get(self.ggml_type, (f"type{self.ggml_type}",))[0]
        return f"GgufTensorInfo({self.name!r}, shape={self.shape}, {t}, offset={self.offset})"


class GgufFile:
    """A GGUF file, memory-mapped.  `metadata[key]` holds Python scalars / str / lists (numeric arrays
    as numpy arrays); `tensor(name)` returns the tensor in numpy's dimension order (GGUF stores the
    innermost dimension first): F32 / F16 as zero-copy views of the mapping, BF16 widened to fp32,
    quantised blocks dequantised to fp32."""

    def __init__(self, path: str):
        self.path = path
        self._f = open(path, "rb")
        try:
            self._mm = mmap.mmap(self._f.fileno(), 0, access=mmap.ACCESS_READ)
        except ValueError as e:                                   # empty file
            self._f.close()
            raise GgufError(f"{path}: not a GGUF file ({e})") from None
        try:
            self._parse()
        except (struct.error, IndexError, UnicodeDecodeError) as e:
            self.close()
            raise GgufError(f"{path}: truncated or corrupt GGUF file ({e})") from None
        except Exception:
            self.close()
            raise

    # -- container ----------------------------------------------------------------------------
    def _parse(self) -> None:
        mm = self._mm
        if len(mm) < 24:
            raise GgufError(f"{self.path}: not a GGUF file (too short)")
        magic, version = struct.unpack_from("<II", mm, 0)
        if magic != GGUF_MAGIC:
            raise GgufError(f"{self.path}: not a GGUF file (magic {magic:#010x})")
        if version not in (2, 3):
            raise GgufError(f"{self.path}: GGUF version {version} is not supported (2 and 3 are)")
        self.version = version
        n_tensors, n_kv = struct.unpack_from("<QQ", mm, 8)
        self._pos = 24
        self.metadata: Dict[str, Any] = {}
        for _ in range(n_kv):
            key = self._string()
            (vtype,) = struct.unpack_from("<I", mm, self._pos)
            self._pos += 4
            self.metadata[key] = self._value(vtype)
        self.alignment = int(self.metadata.get("general.alignment", _DEFAULT_ALIGNMENT))
        if self.alignment <= 0 or self.alignment & (self.alignment - 1):
            raise GgufError(f"{self.path}: general.alignment {self.alignment} is not a power of two")
        infos: List[Tuple[str, Tuple[int, ...], int, int]] = []
        for _ in range(n_tensors):
            name = self._string()
            (n_dims,) = struct.unpack_from("<I", mm, self._pos)
            self._pos += 4
            if n_dims > 8:
                raise GgufError(f"{self.path}: tensor {name!r} claims {n_dims} dimensions")
            dims = struct.unpack_from(f"<{n_dims}Q", mm, self._pos)
            self._pos += 8 * n_dims
            ggml_type, offset = struct.unpack_from("<IQ", mm, self._pos)
            self._pos += 12
            infos.append((name, tuple(int(d) for d in reversed(dims)), int(ggml_type), int(offset)))
        self.data_start = (self._pos + self.alignment - 1) // self.alignment * self.alignment
        self.tensors: Dict[str, GgufTensorInfo] = {}
        for name, shape, ggml_type, offset in infos:
            nbytes = -1
            if ggml_type in GGML_TYPES:
                _, per_block, block_bytes = GGML_TYPES[ggml_type]
                n_elem = int(np.prod(shape, dtype=np.int64)) if shape else 1
                inner = shape[-1] if shape else 1
                if inner % per_block:
                    raise GgufError(f"{self.path}: tensor {name!r}: row length {inner} is not a multiple of the "
                                    f"{GGML_TYPES[ggml_type][0]} block ({per_block})")
                nbytes = n_elem // per_block * block_bytes
                if offset % self.alignment or self.data_start + offset + nbytes > len(mm):
                    raise GgufError(f"{self.path}: tensor {name!r} lies outside the file or is misaligned")
            self.tensors[name] = GgufTensorInfo(name, shape, ggml_type, offset, nbytes)

    def _string(self) -> str:
        (n,) = struct.unpack_from("<Q", self._mm, self._pos)
        self._pos += 8
        if self._pos + n > len(self._mm):
            raise GgufError(f"{self.path}: string of {n} bytes runs past the end of the file")
        s = self._mm[self._pos: self._pos + n].decode("utf-8")
        self._pos += n
        return s

    def _value(self, vtype: int) -> Any:
        if vtype in _SCALAR:
            fmt = _SCALAR[vtype]
            (v,) = struct.unpack_from(fmt, self._mm, self._pos)
            self._pos += struct.calcsize(fmt)
            return v
        if vtype == _T_STR:
            return self._string()
        if vtype == _T_ARR:
            etype, count = struct.unpack_from("<IQ", self._mm, self._pos)
            self._pos += 12
            if etype in _NP_SCALAR:
                dt = np.dtype(_NP_SCALAR[etype])
                if self._pos + count * dt.itemsize > len(self._mm):
                    raise GgufError(f"{self.path}: array of {count} elements runs past the end of the file")
                arr = np.frombuffer(self._mm, dtype=dt, count=count, offset=self._pos).copy()
                self._pos += count * dt.itemsize
                return arr
            if etype == _T_STR:
                return [self._string() for _ in range(count)]
            if etype == _T_ARR:
                return [self._value(_T_ARR) for _ in range(count)]
            raise GgufError(f"{self.path}: unknown array element type {etype}")
        raise GgufError(f"{self.path}: unknown metadata value type {vtype}")

    # -- tensors ------------------------------------------------------------------------------
    def tensor(self, name: str) -> np.ndarray:
        info = self.tensors.get(name)
        if info is None:
            raise KeyError(f"{self.path}: no tensor {name!r}")
        if info.ggml_type not in GGML_TYPES:
            raise GgufError(f"{self.path}: tensor {name!r} has ggml type {info.ggml_type}, which this reader does not "
                            f"decode (supported: {', '.join(v[0] for v in GGML_TYPES.values())})")
        tname = GGML_TYPES[info.ggml_type][0]
        raw = np.frombuffer(self._mm, dtype=np.uint8, count=info.nbytes, offset=self.data_start + info.offset)
        if tname == "F32":
            return raw.view("<f4").reshape(info.shape)
        if tname == "F16":
            return raw.view("<f2").reshape(info.shape)
        if tname == "BF16":
            return (raw.view("<u2").astype(np.uint32) << 16).view(np.float32).reshape(info.shape)
        return _DEQUANT[tname](raw).reshape(info.shape)

    def close(self) -> None:
        mm, self._mm = getattr(self, "_mm", None), None
        if mm is not None:
            try:
                mm.close()
            except BufferError:                                   # zero-copy tensor views are still alive
                pass
        f, self._f = getattr(self, "_f", None), None
        if f is not None:
            f.close()

    def __enter__(self) -> "GgufFile":
        return self

    def __exit__(self, *exc) -> None:
        self.close()


# ggml block formats (ggml-common.h: block_q8_0 / block_q4_0 / block_q4_1), 32 elements per block
def _dq_q8_0(raw: np.ndarray) -> np.ndarray:
    blk = raw.reshape(-1, 34)
    d = blk[:, :2].copy().view("<f2").astype(np.float32)          # [blocks, 1]
    q = blk[:, 2:].view(np.int8).astype(np.float32)
    return (q * d).reshape(-1)


def _nibbles(qs: np.ndarray) -> np.ndarray:
    """16 bytes -> 32 values: the low nibbles are elements 0..15, the high nibbles 16..31."""
    return np.concatenate([qs & 0x0F, qs >> 4], axis=1).astype(np.float32)


def _dq_q4_0(raw: np.ndarray) -> np.ndarray:
    blk = raw.reshape(-1, 18)
    d = blk[:, :2].copy().view("<f2").astype(np.float32)
    return ((_nibbles(blk[:, 2:]) - 8.0) * d).reshape(-1)


def _dq_q4_1(raw: np.ndarray) -> np.ndarray:
    blk = raw.reshape(-1, 20)
    d = blk[:, :2].copy().view("<f2").astype(np.float32)
    m = blk[:, 2:4].copy().view("<f2").astype(np.float32)
    return (_nibbles(blk[:, 4:]) * d + m).reshape(-1)


_DEQUANT = {"Q8_0": _dq_q8_0, "Q4_0": _dq_q4_0, "Q4_1": _dq_q4_1}


# ------------------------------------------------------------------------------ BERT mapping
# llama.cpp's converter (gguf-py tensor_mapping, MODEL_ARCH.BERT) renames BertModel's parameters; this
# is the inverse, GGUF name -> BertModel name.  `{i}` = layer index.
_GLOBAL_NAMES = {
    "token_embd": "embeddings.word_embeddings",
    "position_embd": "embeddings.position_embeddings",
    "token_types": "embeddings.token_type_embeddings",
    "token_embd_norm": "embeddings.LayerNorm",
}
_LAYER_NAMES = {
    "attn_q": "attention.self.query",
    "attn_k": "attention.self.key",
    "attn_v": "attention.self.value",
    "attn_output": "attention.output.dense",
    "attn_output_norm": "attention.output.LayerNorm",
    "ffn_up": "intermediate.dense",
    "ffn_down": "output.dense",
    "layer_output_norm": "output.LayerNorm",
}


def bert_config(g: GgufFile) -> Dict[str, Any]:
    """The hyper-parameters of a BERT GGUF file (`bert.*` keys): layers, hidden, heads, intermediate,
    positions, LayerNorm epsilon, pooling type."""
    md = g.metadata
    arch = md.get("general.architecture")
    if arch != "bert":
        raise GgufError(f"{g.path}: general.architecture is {arch!r}; the encoder runs BERT models "
                        f"(mxbai-embed-large is one)")

    def need(key):
        if key not in md:
            raise GgufError(f"{g.path}: metadata key {key!r} is missing")
        return md[key]
    return {
        "layers": int(need("bert.block_count")),
        "hidden": int(need("bert.embedding_length")),
        "heads": int(need("bert.attention.head_count")),
        "intermediate": int(need("bert.feed_forward_length")),
        "max_positions": int(need("bert.context_length")),
        "eps": float(md.get("bert.attention.layer_norm_epsilon", 1e-12)),
        "pooling": int(md.get("bert.pooling_type", POOLING_NONE)),
        "causal": bool(md.get("bert.attention.causal", False)),
        "name": md.get("general.name", ""),
    }


def bert_state_dict(g: GgufFile) -> Tuple[Dict[str, "Any"], Dict[str, Any]]:
    """(`BertModel` state dict of torch CPU tensors, config).  F16 tensors stay fp16 (the matmul
    operands are stored fp16 on the device anyway), everything else arrives as fp32."""
    import torch
    cfg = bert_config(g)
    sd: Dict[str, Any] = {}

    def put(gguf_name: str, hf_name: str, required: bool = True) -> None:
        for suffix in (".weight", ".bias"):
            if gguf_name + suffix in g.tensors:
                # a private copy: the state dict must outlive the file mapping
                sd[hf_name + suffix] = torch.from_numpy(np.array(g.tensor(gguf_name + suffix)))
            elif required and suffix == ".weight":
                raise GgufError(f"{g.path}: tensor {gguf_name + suffix!r} is missing")

    for gn, hn in _GLOBAL_NAMES.items():
        put(gn, hn)
    for i in range(cfg["layers"]):
        for gn, hn in _LAYER_NAMES.items():
            put(f"blk.{i}.{gn}", f"encoder.layer.{i}.{hn}")
        for gn in ("attn_q", "attn_k", "attn_v", "attn_output", "ffn_up", "ffn_down", "attn_output_norm",
                   "layer_output_norm"):
            if f"blk.{i}.{gn}.bias" not in g.tensors:
                raise GgufError(f"{g.path}: tensor 'blk.{i}.{gn}.bias' is missing (a BERT layer has biases)")
    if f"blk.0.attn_qkv.weight" in g.tensors:
        raise GgufError(f"{g.path}: fused attn_qkv tensors (nomic-bert style) are not a BertModel layout")
    return sd, cfg


def wordpiece_vocab(g: GgufFile) -> Dict[str, int]:
    """`tokenizer.ggml.tokens` of a `tokenizer.ggml.model = "bert"` file -> {piece: id} in BERT's own
    spelling.  The converter's transformation (convert_hf_to_gguf.py, BertModel.set_vocab): a piece in
    square brackets is kept, `##x` becomes `x`, any other piece `x` becomes U+2581 + `x`."""
    md = g.metadata
    model = md.get("tokenizer.ggml.model")
    if model != "bert":
        raise GgufError(f"{g.path}: tokenizer.ggml.model is {model!r}, not a WordPiece ('bert') vocabulary")
    tokens = md.get("tokenizer.ggml.tokens")
    if not isinstance(tokens, list) or not tokens:
        raise GgufError(f"{g.path}: tokenizer.ggml.tokens is missing")
    vocab: Dict[str, int] = {}
    for i, t in enumerate(tokens):
        if t.startswith("[") and t.endswith("]"):
            piece = t
        elif t.startswith("▁"):
            piece = t[1:]
        else:
            piece = "##" + t
        vocab[piece] = i                                          # duplicates: the last id wins, as when a vocab.txt is read
    return vocab


def special_token_ids(g: GgufFile) -> Dict[str, Optional[int]]:
    md = g.metadata

    def get(*keys):
        for k in keys:
            if k in md:
                return int(md[k])
        return None
    return {
        "unk": get("tokenizer.ggml.unknown_token_id"),
        "sep": get("tokenizer.ggml.seperator_token_id", "tokenizer.ggml.separator_token_id",
                   "tokenizer.ggml.eos_token_id"),
        "cls": get("tokenizer.ggml.cls_token_id", "tokenizer.ggml.bos_token_id"),
        "pad": get("tokenizer.ggml.padding_token_id"),
        "mask": get("tokenizer.ggml.mask_token_id"),
    }


# ------------------------------------------------------------------------------ Ollama's store
def ollama_models_dir() -> str:
    return os.environ.get("OLLAMA_MODELS") or os.path.join(os.path.expanduser("~"), ".ollama", "models")


def find_ollama_model(name: str = "mxbai-embed-large", models_dir: Optional[str] = None) -> str:
    """The GGUF blob Ollama serves under `name` ("model", "model:tag", "namespace/model:tag" or
    "host/namespace/model:tag"): manifests/<host>/<namespace>/<model>/<tag> is a JSON manifest whose
    layer of media type application/vnd.ollama.image.model names the blob by digest."""
    root = models_dir or ollama_models_dir()
    tag = "latest"
    rest = name
    if ":" in name.rsplit("/", 1)[-1]:
        rest, tag = name.rsplit(":", 1)
    parts = rest.split("/")
    if len(parts) == 1:
        host, namespace, model = "registry.ollama.ai", "library", parts[0]
    elif len(parts) == 2:
        host, (namespace, model) = "registry.ollama.ai", parts
    elif len(parts) == 3:
        host, namespace, model = parts
    else:
        raise FileNotFoundError(f"cannot parse the model name {name!r}")
    manifest = os.path.join(root, "manifests", host, namespace, model, tag)
    if not os.path.isfile(manifest):
        raise FileNotFoundError(f"Ollama has no manifest for {name!r} under {root} (expected {manifest}); "
                                f"`ollama pull {name}` creates it")
    with open(manifest, encoding="utf-8") as f:
        doc = json.load(f)
    for layer in doc.get("layers", []):
        if layer.get("mediaType") == "application/vnd.ollama.image.model":
            digest = str(layer.get("digest", ""))
            blob = os.path.join(root, "blobs", digest.replace(":", "-"))
            if not os.path.isfile(blob):
                raise FileNotFoundError(f"manifest {manifest} names blob {digest}, which is not in {root}/blobs")
            return blob
    raise FileNotFoundError(f"manifest {manifest} has no model layer")
