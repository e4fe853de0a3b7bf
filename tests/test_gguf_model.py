"""CPU tests of the GGUF reader (semantic-query-engine_b200/gguf_model.py): the file Ollama holds for
the reference's `EMBED_MODEL_NAME` (app/main.py:29) as the encoder's checkpoint.  The reader is pinned on
files written by the `gguf` library's own writer with its own BERT name table and quantisers
(tests/gguf_fixture.py), and on the library's reader; the model a file decodes to is run through the
CPU oracle (oracle/bert_oracle.py) next to the original weights.  No GPU, no compute calls into the
product library."""
import json
import os
import struct

import numpy as np
import pytest
import torch

import sqe_b200
from sqe_b200 import gguf_model as gm
from oracle import bert_oracle as bo

gguf = pytest.importorskip("gguf")
from gguf_fixture import phantom, write_bert_gguf            # noqa: E402

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
VOCAB = json.load(open(os.path.join(GOLDEN, "bert_wordpiece.json"), encoding="utf-8"))["vocab"]


def _weights(seed=5, layers=2, hidden=128, inter=256):
    return bo.random_bert_weights(seed, layers=layers, hidden=hidden, intermediate=inter, vocab=len(VOCAB), max_pos=64)


@pytest.mark.parametrize("ftype", ["f32", "f16", "bf16", "q8_0", "q4_0", "q4_1"])
def test_every_tensor_decodes_to_what_the_library_stored(tmp_path, ftype):
    w = _weights()
    path = str(tmp_path / f"bert-{ftype}.gguf")
    want = write_bert_gguf(path, w, VOCAB, ftype=ftype, heads=4, eps=1e-12)
    with gm.GgufFile(path) as g:
        sd, cfg = gm.bert_state_dict(g)
        assert cfg == {"layers": 2, "hidden": 128, "heads": 4, "intermediate": 256, "max_positions": 64,
                       "eps": pytest.approx(1e-12), "pooling": gm.POOLING_CLS, "causal": False,
                       "name": "mxbai-embed-large-v1"}
        assert g.metadata["general.architecture"] == "bert" and g.version == 3
    assert set(sd) == set(want)
    for k, v in want.items():
        got = sd[k]
        assert tuple(got.shape) == v.shape, k
        if ftype == "f16" and v.ndim == 2 and "position" not in k and "token_type" not in k:
            assert got.dtype == torch.float16, k                      # stored F16 arrives as fp16, bit for bit
        else:
            assert got.dtype == torch.float32, k
        np.testing.assert_array_equal(got.float().numpy(), v, err_msg=k)
    if ftype == "f32":                                              # and an F32 file IS the original model
        for k, v in w.items():
            np.testing.assert_array_equal(sd[k].numpy(), v.numpy(), err_msg=k)


def test_the_decoded_model_is_the_model_for_the_oracle(tmp_path):
    """F32 file -> state dict -> oracle == the original weights through the oracle, bit for bit; the F16
    file (what Ollama ships) stays within fp16 storage error of it."""
    w = _weights(seed=9)
    seqs = [[2, 5, 6, 7, 8, 3], [2, 20, 21, 3], [2] + list(range(5, 60)) + [3]]
    want = bo.bert_embed(w, seqs, heads=4).numpy()
    p32, p16 = str(tmp_path / "a.gguf"), str(tmp_path / "b.gguf")
    write_bert_gguf(p32, w, VOCAB, ftype="f32", heads=4)
    write_bert_gguf(p16, w, VOCAB, ftype="f16", heads=4)
    with gm.GgufFile(p32) as g:
        sd32, cfg = gm.bert_state_dict(g)
    np.testing.assert_array_equal(bo.bert_embed(sd32, seqs, heads=cfg["heads"]).numpy(), want)
    with gm.GgufFile(p16) as g:
        sd16, _ = gm.bert_state_dict(g)
    got16 = bo.bert_embed({k: v.float() for k, v in sd16.items()}, seqs, heads=4).numpy()
    assert np.abs(got16 - want).max() < 2e-2
    cos = (got16 * want).sum(1) / (np.linalg.norm(got16, axis=1) * np.linalg.norm(want, axis=1))
    assert cos.min() > 0.9999


def test_reader_agrees_with_the_library_reader(tmp_path):
    """Tensor directory (names, numpy-order shapes, types, absolute offsets) and metadata against
    `gguf.GGUFReader` on the same file, with a non-default alignment."""
    path = str(tmp_path / "m.gguf")
    write_bert_gguf(path, _weights(), VOCAB, ftype="q8_0", heads=4, alignment=64)
    lib = gguf.GGUFReader(path)
    with gm.GgufFile(path) as g:
        assert g.alignment == 64 and g.data_start == lib.data_offset
        assert [t.name for t in lib.tensors] == list(g.tensors)
        for t in lib.tensors:
            info = g.tensors[t.name]
            assert info.ggml_type == int(t.tensor_type)
            assert info.shape == tuple(int(d) for d in reversed(t.shape))
            assert g.data_start + info.offset == t.data_offset and info.nbytes == t.n_bytes
        for key, field in lib.fields.items():
            if key.startswith("GGUF."):
                continue
            mine = g.metadata[key]
            theirs = field.contents()
            if isinstance(mine, np.ndarray):
                np.testing.assert_array_equal(mine, np.asarray(theirs))
            elif isinstance(mine, float):
                assert mine == pytest.approx(theirs)
            else:
                assert mine == theirs, key


def test_vocabulary_round_trip_and_tokeniser(tmp_path):
    path = str(tmp_path / "m.gguf")
    write_bert_gguf(path, _weights(), VOCAB, ftype="f16", heads=4)
    with gm.GgufFile(path) as g:
        assert g.metadata["tokenizer.ggml.tokens"] == [phantom(t) for t in VOCAB]
        vocab = gm.wordpiece_vocab(g)
        ids = gm.special_token_ids(g)
    assert vocab == {t: i for i, t in enumerate(VOCAB)}
    assert ids["unk"] == VOCAB.index("[UNK]") and ids["sep"] == VOCAB.index("[SEP]") and ids["cls"] == VOCAB.index("[CLS]")
    tok_file = sqe_b200.WordPieceTokenizer.from_gguf(path)
    tok_dict = sqe_b200.WordPieceTokenizer({t: i for i, t in enumerate(VOCAB)})
    cases = json.load(open(os.path.join(GOLDEN, "bert_wordpiece.json"), encoding="utf-8"))["cases"]
    for case in cases:
        text = case["text"] if isinstance(case, dict) else case[0]
        assert tok_file.encode(text) == tok_dict.encode(text)
        assert tok_file.encode(text) == bo.encode_text(text, vocab)


def test_ollama_store_lookup(tmp_path):
    root = tmp_path / "models"
    blob_dir = root / "blobs"
    blob_dir.mkdir(parents=True)
    digest = "sha256:" + "ab" * 32
    blob = blob_dir / digest.replace(":", "-")
    write_bert_gguf(str(blob), _weights(layers=1), VOCAB, ftype="f16", heads=4)
    man = root / "manifests" / "registry.ollama.ai" / "library" / "mxbai-embed-large"
    man.mkdir(parents=True)
    doc = {"schemaVersion": 2, "mediaType": "application/vnd.docker.distribution.manifest.v2+json",
           "config": {"mediaType": "application/vnd.docker.container.image.v1+json", "digest": "sha256:" + "00" * 32},
           "layers": [{"mediaType": "application/vnd.ollama.image.license", "digest": "sha256:" + "11" * 32},
                      {"mediaType": "application/vnd.ollama.image.model", "digest": digest, "size": blob.stat().st_size},
                      {"mediaType": "application/vnd.ollama.image.params", "digest": "sha256:" + "22" * 32}]}
    (man / "latest").write_text(json.dumps(doc))
    (man / "335m").write_text(json.dumps(doc))
    assert gm.find_ollama_model("mxbai-embed-large", str(root)) == str(blob)
    assert gm.find_ollama_model("mxbai-embed-large:335m", str(root)) == str(blob)
    assert gm.find_ollama_model("library/mxbai-embed-large:latest", str(root)) == str(blob)
    assert gm.find_ollama_model("registry.ollama.ai/library/mxbai-embed-large", str(root)) == str(blob)
    os.environ["OLLAMA_MODELS"] = str(root)
    try:
        assert gm.find_ollama_model() == str(blob)
    finally:
        del os.environ["OLLAMA_MODELS"]
    with pytest.raises(FileNotFoundError, match="ollama pull"):
        gm.find_ollama_model("nomic-embed-text", str(root))
    blob.unlink()
    with pytest.raises(FileNotFoundError, match="not in"):
        gm.find_ollama_model("mxbai-embed-large", str(root))
    with gm.GgufFile.__new__(gm.GgufFile) as _:                     # close() on a never-opened object is harmless
        pass


def test_malformed_and_foreign_files_are_refused(tmp_path):
    good = str(tmp_path / "good.gguf")
    write_bert_gguf(good, _weights(layers=1), VOCAB, ftype="f16", heads=4)
    raw = open(good, "rb").read()

    def variant(name, data):
        p = str(tmp_path / name)
        with open(p, "wb") as f:
            f.write(data)
        return p
    with pytest.raises(gm.GgufError, match="not a GGUF file"):
        gm.GgufFile(variant("magic.gguf", b"GGML" + raw[4:]))
    with pytest.raises(gm.GgufError, match="not a GGUF file"):
        gm.GgufFile(variant("empty.gguf", b""))
    with pytest.raises(gm.GgufError, match="version 1"):
        gm.GgufFile(variant("v1.gguf", raw[:4] + struct.pack("<I", 1) + raw[8:]))
    with pytest.raises(gm.GgufError):
        gm.GgufFile(variant("cut_header.gguf", raw[:300]))
    with pytest.raises(gm.GgufError, match="outside the file"):
        gm.GgufFile(variant("cut_data.gguf", raw[: len(raw) - 4096]))
    # another architecture / another tokeniser / a tensor type the reader does not decode
    llama = str(tmp_path / "llama.gguf")
    write_bert_gguf(llama, _weights(layers=1), VOCAB, ftype="f16", heads=4, arch="nomic-bert")
    with gm.GgufFile(llama) as g, pytest.raises(gm.GgufError, match="architecture"):
        gm.bert_state_dict(g)
    w = gguf.GGUFWriter(str(tmp_path / "q6.gguf"), "bert")
    w.add_tokenizer_model("gpt2")
    w.add_tensor("token_embd.weight", np.zeros((2, 210), np.uint8),             # two rows of one 256-element Q6_K block
                 raw_dtype=gguf.GGMLQuantizationType.Q6_K)
    w.write_header_to_file(); w.write_kv_data_to_file(); w.write_tensors_to_file(); w.close()
    with gm.GgufFile(str(tmp_path / "q6.gguf")) as g:
        with pytest.raises(gm.GgufError, match="does not decode"):
            g.tensor("token_embd.weight")
        with pytest.raises(KeyError):
            g.tensor("nope")
        with pytest.raises(gm.GgufError, match="WordPiece"):
            gm.wordpiece_vocab(g)
        with pytest.raises(gm.GgufError, match="block_count"):
            gm.bert_state_dict(g)
    # a layer tensor missing
    w1 = _weights(layers=1)
    del w1["encoder.layer.0.output.dense.bias"]
    p = str(tmp_path / "nobias.gguf")
    write_bert_gguf(p, w1, VOCAB, ftype="f32", heads=4)
    with gm.GgufFile(p) as g, pytest.raises(gm.GgufError, match="ffn_down.bias"):
        gm.bert_state_dict(g)


def test_encoder_entry_points_refuse_what_the_kernels_are_not_built_for(tmp_path):
    """`EncoderWeights.from_gguf` checks the file's hyper-parameters BEFORE touching a device: geometry,
    pooling type; and like every product entry point it has no CPU path."""
    p = str(tmp_path / "small.gguf")
    write_bert_gguf(p, _weights(layers=1), VOCAB, ftype="f16", heads=4)
    with pytest.raises(ValueError, match="4 heads x hidden 128"):
        sqe_b200.EncoderWeights.from_gguf(p, device="cpu")
    w = bo.random_bert_weights(1, layers=1, vocab=len(VOCAB))
    p2 = str(tmp_path / "mean.gguf")
    write_bert_gguf(p2, w, VOCAB, ftype="f16", pooling=gm.POOLING_MEAN)
    with pytest.raises(ValueError, match="pooling type 1"):
        sqe_b200.EncoderWeights.from_gguf(p2, device="cpu")
    p3 = str(tmp_path / "cls.gguf")
    write_bert_gguf(p3, w, VOCAB, ftype="f16")
    with pytest.raises(RuntimeError, match="no CPU path"):
        sqe_b200.EncoderWeights.from_gguf(p3, device="cpu")
