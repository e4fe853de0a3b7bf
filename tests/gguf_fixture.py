"""Test helper: write a BERT GGUF file the way llama.cpp's converter does, with the `gguf` library's
OWN writer, name table and quantisers (library = the published implementation the product's reader
in gguf_model.py is pinned on; nothing here is product code).

What the converter does for a BertModel (convert_hf_to_gguf.py, class BertModel): tensor names through
`gguf.get_tensor_name_map(MODEL_ARCH.BERT, n_layers)`; 1-D tensors and the position / token-type tables
stay F32, 2-D weights take the file type; the vocabulary is rewritten (pieces in brackets kept, `##x` ->
`x`, other pieces `x` -> U+2581 + `x`); pooling type, LayerNorm epsilon, non-causal attention in the
`bert.*` keys."""
from typing import Dict, Sequence

import numpy as np


def phantom(tok: str) -> str:
    if tok.startswith("[") and tok.endswith("]"):
        return tok
    if tok.startswith("##"):
        return tok[2:]
    return "▁" + tok


def write_bert_gguf(path: str, sd: Dict[str, "np.ndarray"], vocab: Sequence[str], *, ftype: str = "f16",
                    heads: int = 16, eps: float = 1e-12, pooling: int = 2, arch: str = "bert",
                    name: str = "mxbai-embed-large-v1", alignment: int = 32) -> Dict[str, np.ndarray]:
    """`sd`: BertModel names -> torch tensors / numpy arrays (fp32).  Returns what a loader should see:
    BertModel name -> the fp32 values the file holds (after the file type's rounding)."""
    import gguf
    import torch
    sd = {k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)).astype(np.float32)
          for k, v in sd.items()}
    layers = 0
    while f"encoder.layer.{layers}.attention.self.query.weight" in sd:
        layers += 1
    hidden = sd["embeddings.word_embeddings.weight"].shape[1]
    inter = sd["encoder.layer.0.intermediate.dense.weight"].shape[0]
    w = gguf.GGUFWriter(path, arch)
    if alignment != 32:
        w.add_custom_alignment(alignment)
    w.add_name(name)
    w.add_block_count(layers)
    w.add_context_length(sd["embeddings.position_embeddings.weight"].shape[0])
    w.add_embedding_length(hidden)
    w.add_feed_forward_length(inter)
    w.add_head_count(heads)
    w.add_layer_norm_eps(eps)
    w.add_causal_attention(False)
    w.add_pooling_type(gguf.PoolingType(pooling))
    w.add_file_type({"f32": 0, "f16": 1, "q8_0": 7, "q4_0": 2, "q4_1": 3, "bf16": 32}[ftype])
    w.add_tokenizer_model("bert")
    w.add_token_list([phantom(t) for t in vocab])
    w.add_token_types([3 if (t.startswith("[") and t.endswith("]")) else 1 for t in vocab])
    w.add_token_type_count(2)
    ids = {t: i for i, t in enumerate(vocab)}
    for tok, add in (("[UNK]", w.add_unk_token_id), ("[SEP]", w.add_sep_token_id), ("[PAD]", w.add_pad_token_id),
                     ("[MASK]", w.add_mask_token_id), ("[CLS]", w.add_bos_token_id)):
        if tok in ids:
            add(ids[tok])
    tmap = gguf.get_tensor_name_map(gguf.MODEL_ARCH.BERT, layers)
    qt = {"q8_0": gguf.GGMLQuantizationType.Q8_0, "q4_0": gguf.GGMLQuantizationType.Q4_0,
          "q4_1": gguf.GGMLQuantizationType.Q4_1}
    seen: Dict[str, np.ndarray] = {}
    for hf_name, arr in sd.items():
        if hf_name.startswith("pooler.") or hf_name.endswith("position_ids"):
            continue                                              # the converter drops them
        new = tmap.get_name(hf_name, try_suffixes=(".weight", ".bias"))
        assert new is not None, hf_name
        keep_f32 = arr.ndim <= 1 or new.startswith(("position_embd", "token_types"))
        if keep_f32 or ftype == "f32":
            w.add_tensor(new, arr)
            seen[hf_name] = arr
        elif ftype == "f16":
            h = arr.astype(np.float16)
            w.add_tensor(new, h)
            seen[hf_name] = h.astype(np.float32)
        elif ftype == "bf16":
            packed = gguf.quants.quantize(arr, gguf.GGMLQuantizationType.BF16)
            w.add_tensor(new, packed, raw_dtype=gguf.GGMLQuantizationType.BF16)
            seen[hf_name] = gguf.quants.dequantize(packed, gguf.GGMLQuantizationType.BF16).reshape(arr.shape)
        else:
            packed = gguf.quants.quantize(arr, qt[ftype])
            w.add_tensor(new, packed, raw_dtype=qt[ftype])
            seen[hf_name] = gguf.quants.dequantize(packed, qt[ftype]).reshape(arr.shape)
    w.write_header_to_file()
    w.write_kv_data_to_file()
    w.write_tensors_to_file()
    w.close()
    return seen
