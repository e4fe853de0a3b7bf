"""Generate tests/golden/bert_tiny.npz and tests/golden/bert_wordpiece.json: outputs of the
PUBLISHED implementations the encoder oracle restates (oracle/bert_oracle.py) --
`transformers.BertModel` for the arithmetic and the `tokenizers` library's BERT pipeline for the
tokeniser.  Run in the build container (both libraries are in the image):

    python oracle/make_bert_golden.py

TEST INFRASTRUCTURE ONLY.  The fixtures are small on purpose: a 2-layer, hidden-64 BertModel with
seeded random weights (weights, token ids and last hidden states are all stored), and a synthetic
WordPiece vocabulary with hostile strings.
"""
import json
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")

VOCAB = ["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]", "the", "quick", "brown", "fox", "##es", "jump", "##s",
         "##ed", "##ing", "over", "lazy", "dog", ",", ".", "!", "?", "-", "(", ")", "'", "a", "##b", "##c", "un",
         "##aff", "##able", "e", "cafe", "naive", "resume", "中", "文", "embed", "##ding", "vector", "##s", "search",
         "query", "1024", "10", "##24", "mxbai", "large", "semantic", "engine", "b", "##200", "hello", "world",
         "new", "york", "o", "reilly", "$", "5", "##0", "%", "u", "s", "##a", "co", "##operate", "re", "##sum"]
TEXTS = [
    "The quick brown foxes jumped over the lazy dog!",
    "unaffable, abc. É 中文 xyz",
    "  ",
    "",
    "a" * 150,
    "ab c� d\te\nf  g\x00h​i",
    "Café naïve résumé -- co-operate (O'Reilly) $50%?!",
    "Hello,World!Hello , world .",
    "embedding vectors search query 1024 mxbai large semantic engine b200",
    "New York　U.S.A",
    "中文embedding文",
    "jumping jumps jumped jumpss",
]


def bert_tiny():
    from transformers import BertConfig, BertModel
    torch.manual_seed(20260)
    cfg = BertConfig(hidden_size=64, num_hidden_layers=2, num_attention_heads=4, intermediate_size=128,
                     vocab_size=97, max_position_embeddings=40, hidden_act="gelu", layer_norm_eps=1e-12)
    m = BertModel(cfg, add_pooling_layer=False).eval()
    with torch.no_grad():
        for name, p in m.named_parameters():          # spread the weights: peaky softmax rows, live LayerNorms
            if "query.weight" in name or "key.weight" in name:
                p.normal_(0.0, 0.35)
            elif "LayerNorm.weight" in name:
                p.copy_(1.0 + 0.1 * torch.randn_like(p))
            elif p.dim() == 1:
                p.normal_(0.0, 0.1)
            elif "embeddings" in name:
                p.normal_(0.0, 0.5)
            else:
                p.normal_(0.0, 0.12)
    g = torch.Generator().manual_seed(7)
    lens = [1, 2, 7, 16, 33, 40]
    seqs = [torch.randint(0, 97, (n,), generator=g) for n in lens]
    # padded batch + mask through the library, exactly as a server would run it
    L = max(lens)
    ids = torch.zeros(len(lens), L, dtype=torch.long)
    mask = torch.zeros(len(lens), L, dtype=torch.long)
    for i, s in enumerate(seqs):
        ids[i, : len(s)] = s
        mask[i, : len(s)] = 1
    with torch.no_grad():
        out = m(input_ids=ids, attention_mask=mask).last_hidden_state
    arrays = {"w/" + k: v.numpy() for k, v in m.state_dict().items() if "position_ids" not in k}
    for i, s in enumerate(seqs):
        arrays[f"ids/{i}"] = s.numpy().astype(np.int32)
        arrays[f"hidden/{i}"] = out[i, : len(s)].numpy()
    np.savez_compressed(os.path.join(GOLDEN, "bert_tiny.npz"), **arrays)
    print("bert_tiny.npz:", len(arrays), "arrays")


def wordpiece():
    from tokenizers import Tokenizer
    from tokenizers.models import WordPiece
    from tokenizers.normalizers import BertNormalizer
    from tokenizers.pre_tokenizers import BertPreTokenizer
    vocab = {t: i for i, t in enumerate(VOCAB)}
    tk = Tokenizer(WordPiece(vocab, unk_token="[UNK]", max_input_chars_per_word=100))
    tk.normalizer = BertNormalizer(clean_text=True, handle_chinese_chars=True, strip_accents=None, lowercase=True)
    tk.pre_tokenizer = BertPreTokenizer()
    cases = [{"text": t, "ids": tk.encode(t).ids} for t in TEXTS]
    with open(os.path.join(GOLDEN, "bert_wordpiece.json"), "w", encoding="utf-8") as f:
        json.dump({"vocab": VOCAB, "cases": cases}, f, ensure_ascii=True, indent=1)
    print("bert_wordpiece.json:", len(cases), "cases")


if __name__ == "__main__":
    bert_tiny()
    wordpiece()
