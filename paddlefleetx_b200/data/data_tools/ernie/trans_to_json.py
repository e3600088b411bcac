"""Raw text -> jsonl for the ERNIE corpus pipeline (reference ppfleetx/data/data_tools/ernie/preprocess/trans_to_json.py); the converter is
shared with GPT: ``data_tools/gpt/raw_trans_to_json.py``."""
from ..gpt.raw_trans_to_json import get_args, main, merge_file, raw_text_to_json, shuffle_file  # noqa: F401

if __name__ == "__main__":
    main()
