from .gpt_dataset import GPTDataset, LM_Eval_Dataset, Lambada_Eval_Dataset, SyntheticGPTDataset  # noqa: F401
