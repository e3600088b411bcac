"""Tokenizer stack: WordPiece (ERNIE), sentencepiece (T5 / DeBERTa-v2 / Chinese GPT) and the shared padding / truncation / persistence
machinery.  Vocabularies are built in the test (sentencepiece trains a ~60-piece model in milliseconds); nothing is downloaded."""
import io
import os

import pytest
import torch

CORPUS = ["the quick brown fox jumps over the lazy dog", "hello world this is a tiny corpus", "sentencepiece trains offline",
          "another line of words for the model", "a photo of a small red bird on a branch"] * 20


def _train_spm(path, **kw):
    import sentencepiece as spm

    buf = io.BytesIO()
    spm.SentencePieceTrainer.train(sentence_iterator=iter(CORPUS), model_writer=buf, vocab_size=70, model_type="unigram", hard_vocab_limit=False,
                                   minloglevel=2, **kw)
    with open(path, "wb") as f:
        f.write(buf.getvalue())


@pytest.fixture(scope="module")
def t5_dir(tmp_path_factory):
    d = tmp_path_factory.mktemp("t5")
    _train_spm(d / "spiece.model", pad_id=0, eos_id=1, unk_id=2, bos_id=-1, pad_piece="<pad>", eos_piece="</s>", unk_piece="<unk>")
    return str(d)


@pytest.fixture(scope="module")
def deberta_dir(tmp_path_factory):
    d = tmp_path_factory.mktemp("deberta")
    _train_spm(d / "spm.model", pad_id=0, bos_id=1, eos_id=2, unk_id=3, pad_piece="[PAD]", bos_piece="[CLS]", eos_piece="[SEP]", unk_piece="[UNK]")
    return str(d)


@pytest.fixture(scope="module")
def ernie_dir(tmp_path_factory):
    d = tmp_path_factory.mktemp("ernie")
    words = ["[PAD]", "[CLS]", "[SEP]", "[MASK]", "[UNK]", "the", "quick", "brown", "fox", "jump", "##s", "##ing", "over", "lazy", "dog", ",", ".", "!",
             "un", "##believ", "##able", "cafe", "中", "国", "人", "hello", "world"]
    (d / "vocab.txt").write_text("\n".join(words) + "\n", encoding="utf-8")
    return str(d)


def test_ernie_wordpiece_template_and_roundtrip(ernie_dir, tmp_path):
    from paddlefleetx_b200.data.tokenizers import ErnieTokenizer, get_ernie_tokenizer

    tok = ErnieTokenizer.from_pretrained(ernie_dir)
    assert tok.tokenize("The quick fox jumps, unbelievable!") == ["the", "quick", "fox", "jump", "##s", ",", "un", "##believ", "##able", "!"]
    assert tok.tokenize("Café 中国人 zzz") == ["cafe", "中", "国", "人", "[UNK]"]          # accents stripped, CJK isolated, OOV -> [UNK]
    assert tok.tokenize("hello [MASK] world") == ["hello", "[MASK]", "world"]              # special tokens survive the basic tokenizer
    enc = tok("the fox", "lazy dog")
    v = tok.vocab
    assert enc["input_ids"] == [v["[CLS]"], v["the"], v["fox"], v["[SEP]"], v["lazy"], v["dog"], v["[SEP]"]]
    assert enc["token_type_ids"] == [0, 0, 0, 0, 1, 1, 1] and enc["attention_mask"] == [1] * 7
    assert tok.decode(enc["input_ids"], skip_special_tokens=True) == "the fox lazy dog"
    assert tok.decode(tok.encode("jumps over", add_special_tokens=False)) == "jumps over"
    assert tok.get_special_tokens_mask([5, 6], [7]) == [1, 0, 0, 1, 0, 1] and tok.num_special_tokens_to_add(pair=True) == 3
    # batch: padding to the longest, truncation (longest_first trims the longer member of a pair), tensors
    b = tok(["the quick brown fox", "dog"], padding=True, return_tensors="pt")
    assert b.input_ids.shape == (2, 6) and b.attention_mask[1].tolist() == [1, 1, 1, 0, 0, 0] and b.input_ids[1, -1].item() == tok.pad_token_id
    t = tok("the quick brown fox jumps", "dog", truncation=True, max_length=7)
    assert len(t["input_ids"]) == 7 and t["input_ids"][-2] == v["dog"]
    o = tok("the quick brown fox", truncation="only_first", max_length=4, return_overflowing_tokens=True, stride=1)
    assert o["num_truncated_tokens"] == 2 and o["overflowing_tokens"] == [v["quick"], v["brown"], v["fox"]]
    m = tok(["the", "the quick"], padding="max_length", max_length=6, pad_to_multiple_of=8)
    assert all(len(r) == 8 for r in m["input_ids"])
    # added tokens get fresh ids past the vocabulary and are never split
    assert tok.add_tokens(["<ent>"]) == 1 and tok.convert_tokens_to_ids("<ent>") == tok.vocab_size
    assert tok.tokenize("the <ent> dog") == ["the", "<ent>", "dog"] and len(tok) == tok.vocab_size + 1
    # persistence
    tok.save_pretrained(str(tmp_path))
    again = ErnieTokenizer.from_pretrained(str(tmp_path))
    assert again.tokenize("the <ent> dog") == ["the", "<ent>", "dog"] and again.vocab == tok.vocab
    assert get_ernie_tokenizer(ernie_dir) is get_ernie_tokenizer(ernie_dir)
    with pytest.raises(FileNotFoundError, match="offline"):
        ErnieTokenizer.from_pretrained("ernie-1.0-not-here")


def test_t5_tokenizer_sentinels_eos_and_batch(t5_dir, tmp_path):
    from paddlefleetx_b200.data.tokenizers import T5Tokenizer, t5_tokenize

    tok = T5Tokenizer.from_pretrained(t5_dir, extra_ids=10)
    base = tok.sp_model.get_piece_size()
    assert tok.vocab_size == base + 10 and len(tok.additional_special_tokens) == 10
    assert tok.convert_tokens_to_ids("<extra_id_0>") == tok.vocab_size - 1 and tok.convert_ids_to_tokens(base) == "<extra_id_9>"
    ids = tok("the quick <extra_id_0> fox")["input_ids"]
    assert ids[-1] == tok.eos_token_id == 1 and (tok.vocab_size - 1) in ids
    assert tok.decode(ids, skip_special_tokens=True) == "the quick fox"
    pair = tok("the fox", "lazy dog")
    assert pair["input_ids"].count(tok.eos_token_id) == 2 and "token_type_ids" not in pair
    assert tok.build_inputs_with_special_tokens([5, 1]) == [5, 1]                           # no second </s> when one is present
    ids_t, mask_t = t5_tokenize(["a photo of a small red bird", "dog"], tok, max_length=6)
    assert ids_t.shape == mask_t.shape and ids_t.shape[1] <= 6 and torch.is_tensor(ids_t) and ids_t[1, -1].item() == tok.pad_token_id == 0
    assert mask_t[1].sum().item() == len(tok("dog")["input_ids"])
    tok.save_pretrained(str(tmp_path))
    again = T5Tokenizer.from_pretrained(str(tmp_path))
    assert os.path.isfile(tmp_path / "spiece.model") and again.vocab_size == tok.vocab_size and again.encode("hello world") == tok.encode("hello world")
    with pytest.raises(ValueError):
        T5Tokenizer(os.path.join(t5_dir, "spiece.model"), extra_ids=3, additional_special_tokens=["<extra_id_0>"])


def test_debertav2_tokenizer_template_and_mask_token(deberta_dir):
    from paddlefleetx_b200.data.tokenizers import DebertaV2Tokenizer, debertav2_tokenize

    tok = DebertaV2Tokenizer.from_pretrained(deberta_dir)
    n = tok._tokenizer.spm.get_piece_size()
    assert (tok.pad_token_id, tok.cls_token_id, tok.sep_token_id, tok.unk_token_id) == (0, 1, 2, 3) and tok.mask_token_id == n == tok.vocab_size - 1
    enc = tok("hello world", "lazy dog")
    ids = enc["input_ids"]
    assert ids[0] == 1 and ids[-1] == 2 and ids.count(2) == 2
    first = ids.index(2) + 1
    assert enc["token_type_ids"] == [0] * first + [1] * (len(ids) - first)
    assert tok.decode(tok("hello world")["input_ids"], skip_special_tokens=True) == "hello world"
    assert "[MASK]" in tok.tokenize("hello [MASK] world")
    ids_t, mask_t = debertav2_tokenize(["a photo of a small red bird", "dog"], tok)
    assert ids_t.shape == mask_t.shape and (ids_t[:, 0] == 1).all() and ids_t[1, -1].item() == 0
    punct = DebertaV2Tokenizer.from_pretrained(deberta_dir, split_by_punct=True)
    assert "".join(punct.tokenize("dog,fox")).replace("▁", "").count(",") == 1


def test_gpt_chinese_tokenizer_space_newline_roundtrip(tmp_path):
    from paddlefleetx_b200.data.tokenizers import GPTChineseTokenizer

    _train_spm(tmp_path / "sentencepiece.model", user_defined_symbols=["▂", "▃", "<eod>", "<bod>"])
    tok = GPTChineseTokenizer.from_pretrained(str(tmp_path))
    ids = tok.encode("hello world\nthe fox")
    assert tok.eol_token_id in ids and tok.decode(ids) == "hello world\nthe fox"
    assert tok.eos_token_id == tok.convert_tokens_to_ids("<eod>") and len(tok) == tok.vocab_size
    out = tok(["hello", "hello world"], padding=True)
    assert len(out["input_ids"][0]) == len(out["input_ids"][1]) and out["attention_mask"][0][0] == 0     # left padding


def test_pad_collates_feature_lists_and_numpy(ernie_dir):
    import numpy as np

    from paddlefleetx_b200.data.tokenizers import BatchEncoding, ErnieTokenizer

    tok = ErnieTokenizer.from_pretrained(ernie_dir, padding_side="left")
    feats = [tok("the fox"), tok("the quick brown fox")]
    batch = tok.pad(feats, return_tensors="np")
    assert isinstance(batch, BatchEncoding) and isinstance(batch["input_ids"], np.ndarray) and batch["input_ids"].shape == (2, 6)
    assert batch["attention_mask"][0].tolist() == [0, 0, 1, 1, 1, 1] and batch["token_type_ids"].shape == (2, 6)
    with pytest.raises(ValueError, match="rectangular"):
        tok(["the fox", "the quick brown fox"], return_tensors="pt")
