"""Serve an exported GPT generation model (reference projects/gpt/inference.py:42-66) — same driver as tasks/gpt/inference.py."""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))

from tasks.gpt.inference import main  # noqa: E402

if __name__ == "__main__":
    main()
