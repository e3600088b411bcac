"""Corpus tools: Chinese whole-word marking, sentence splitters, the Converter that turns jsonl lines into id lists (reference
data_tools/gpt/preprocess_data.py, ernie/preprocess/*)."""
import json
import os

import numpy as np
import pytest

from paddlefleetx_b200.data.data_tools.gpt import preprocess_data as P


def test_whole_word_marks_follow_the_segmentation():
    tokens = ["通", "过", "利", "用", "me", "##rc", "##er", "核", "，", "将", "样", "本"]
    words = ["通过", "利用", "mercer", "核", "，", "将", "样本"]
    assert P.get_whole_word_mask_tokens(tokens, words) == ["通", "##过", "利", "##用", "me", "##rc", "##er", "核", "，", "将", "样", "##本"]
    # longest match first, capped at max_word_length; unknown characters stay single
    assert P.get_whole_word_mask_tokens(list("中华人民共和国"), ["中华", "中华人民", "共和国"]) == ["中", "##华", "##人", "##民", "共", "##和", "##国"]
    assert P.get_whole_word_mask_tokens(list("中华人民"), ["中华人民"], max_word_length=2) == list("中华人民")
    assert P.get_whole_word_mask_tokens([], ["x"]) == []


def test_splitters():
    assert P.IdentitySplitter().tokenize("a. b.") == ("a. b.",)
    assert P.NewlineSplitter().tokenize("a\nb") == ["a", "b"]
    assert P._ChineseSplitter().tokenize("第一句。第二句！") == ["第一句。", "第二句！"]
    assert P._ChineseSplitter().tokenize("第一句\n第二句") == ["第一句", "第二句"]
    assert list(P._RegexSplitter(False).tokenize("One. Two? Three")) == ["One.", "Two?", "Three"]


@pytest.fixture()
def ernie_vocab(tmp_path):
    words = ["[PAD]", "[CLS]", "[SEP]", "[MASK]", "[UNK]", "通", "过", "利", "用", "核", "##过", "##用", "，", "me", "##rc", "##er", "hello", "world"]
    d = tmp_path / "ernie"
    d.mkdir()
    (d / "vocab.txt").write_text("\n".join(words) + "\n", encoding="utf-8")
    return str(d), {w: i for i, w in enumerate(words)}


def test_converter_whole_word_ids_and_corpus_files(tmp_path, ernie_vocab):
    vocab_dir, vocab = ernie_vocab
    args = P.get_args(["--input_path", "x", "--output_prefix", "y", "--tokenizer_name", "ErnieTokenizer", "--model_name", vocab_dir, "--chinese",
                       "--cn_whole_word_segment", "--cn_splited", "--cn_split_dimer", " ", "--split_sentences", "--append_eos"])
    conv = P.Converter(args)
    conv.initializer()
    doc, nbytes = conv.encode(json.dumps({"text": "通过 利用 mercer 核\n通过"}, ensure_ascii=False))
    want_first = [vocab[t] for t in ["通", "##过", "利", "##用", "me", "##rc", "##er", "核"]]
    assert doc[0] == want_first and doc[1] == [vocab["通"], vocab["##过"], vocab["[SEP]"]] and nbytes > 0
    # without whole-word segmentation the continuation characters keep their plain ids
    plain = P.get_args(["--input_path", "x", "--output_prefix", "y", "--tokenizer_name", "ErnieTokenizer", "--model_name", vocab_dir, "--chinese"])
    conv2 = P.Converter(plain)
    conv2.initializer()
    doc2, _ = conv2.encode(json.dumps({"text": "通过"}, ensure_ascii=False))
    assert doc2 == [[vocab["通"], vocab["过"]]]
    # end to end: jsonl -> _ids.npy / _idx.npz with sentence and document boundaries
    src = tmp_path / "corpus.jsonl"
    src.write_text("\n".join(json.dumps({"text": t}, ensure_ascii=False) for t in ["通过 利用\n核", "hello world"]) + "\n", encoding="utf-8")
    out = str(tmp_path / "out")
    P.main(["--input_path", str(src), "--output_prefix", out, "--tokenizer_name", "ErnieTokenizer", "--model_name", vocab_dir, "--chinese",
            "--cn_whole_word_segment", "--cn_splited", "--split_sentences"])
    idx = np.load(out + "_idx.npz")
    ids = np.load(out + "_ids.npy")
    assert idx["lens"].tolist() == [4, 1, 2] and idx["docs"].tolist() == [0, 2, 3]
    assert ids[:4].tolist() == [vocab["通"], vocab["##过"], vocab["利"], vocab["##用"]]


def test_segmenter_factories_are_lazy():
    jieba = pytest.importorskip("jieba", reason="jieba not installed")
    assert "".join(P.jieba_segmentation_fn()("通过利用")) == "通过利用" and jieba is not None


def test_ernie_front_ends_share_the_converter():
    from paddlefleetx_b200.data.data_tools.ernie import create_pretraining_data as C

    for name in ("get_args", "lexical_analysis_fn", "chinese_segmentation_fn", "jieba_segmentation_fn", "get_whole_word_mask_tokens", "IdentitySplitter",
                 "NewlineSplitter", "Converter"):
        assert getattr(C, name) is getattr(P, name) or name == "get_args"


def test_raw_text_tools(tmp_path):
    from paddlefleetx_b200.data.data_tools.gpt import raw_trans_to_json as R

    raw = tmp_path / "a.txt"
    raw.write_text("first document line one\nline two\n\nsecond document is here\n\nshort\n", encoding="utf-8")
    nbytes, out = R.raw_text_to_json(str(raw), doc_spliter="", min_doc_length=10)
    docs = [json.loads(l)["text"] for l in open(out, encoding="utf-8")]
    assert nbytes == os.path.getsize(raw) and docs == ["first document line one\nline two", "second document is here"]
    assert R.raw_text_to_json(str(tmp_path / "missing.txt")) == (0, None)
    other = tmp_path / "b.txt.jsonl"
    other.write_text(json.dumps({"text": "third document, long enough"}) + "\n", encoding="utf-8")
    merged = R.merge_file([out, str(other), None], str(tmp_path / "all"))
    assert merged.endswith("all.jsonl") and not os.path.exists(out) and len(open(merged).readlines()) == 3
    before = open(merged).readlines()
    R.shuffle_file(merged)
    after = open(merged).readlines()
    assert sorted(before) == sorted(after)
    with pytest.raises(ValueError):
        R.shuffle_file(str(tmp_path / "nope.jsonl"))


def test_words_segmentation_flows(tmp_path):
    from paddlefleetx_b200.data.data_tools.ernie import words_segmentation as W

    seg = lambda line: line.split("|") if "|" in line else list(line)      # noqa: E731 - a deterministic stand-in for jieba
    assert W.split_and_segment("你好。。  再见!ok", seg) == "你 好 。\n再 见 !\no k\n"
    src = tmp_path / "docs.jsonl"
    src.write_text("\n".join(json.dumps({"text": t}, ensure_ascii=False) for t in ["你好。再见", "第二篇"]) + "\n", encoding="utf-8")
    assert list(W.read_jsonl(str(src))) == ["你好。再见", "第二篇"]
    wudao = tmp_path / "shard.json"
    wudao.write_text(json.dumps([{"content": "一"}, {"content": "二"}], ensure_ascii=False), encoding="utf-8")
    assert list(W.read_wudao(str(wudao))) == ["一", "二"]
    out_dir = tmp_path / "seg"
    out_dir.mkdir()
    nbytes, _ = W.text_to_text(str(src), str(out_dir), "jsonl", list)
    text = open(out_dir / "docs.jsonl", encoding="utf-8").read()
    assert nbytes > 0 and text == "你 好 。\n再 见\n\n第 二 篇\n\n"
    assert W.text_to_text(str(src), str(out_dir), "jsonl", list) == (0, None)         # existing outputs are kept
    # jsonl -> jsonl flow through the CLI (character fallback or jieba, whichever is installed: the characters are preserved either way)
    dst = tmp_path / "docs_seg.jsonl"
    W.main(["--input_path", str(src), "--output_path", str(dst), "--cn_split_dimer", "|"])
    rows = [json.loads(l)["text"] for l in open(dst, encoding="utf-8")]
    assert [r.replace("|", "") for r in rows] == ["你好。再见", "第二篇"] and "|" in rows[0]
