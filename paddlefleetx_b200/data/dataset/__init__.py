from .gpt_dataset import GPTDataset, LM_Eval_Dataset, Lambada_Eval_Dataset, SyntheticGPTDataset  # noqa: F401
from .glue_dataset import CoLA, MNLI, MRPC, QNLI, QQP, RTE, SST2, STSB, WNLI, GlueDataset  # noqa: F401,E402
from .ernie.ernie_dataset import ErnieDataset, ErnieSeqClsDataset, SyntheticErnieDataset  # noqa: F401,E402
from .vision_dataset import CIFAR10, ContrativeLearningDataset, GeneralClsDataset, ImageFolder, SyntheticImageDataset  # noqa: F401,E402
from .multimodal_dataset import ImagenDataset, SyntheticImagenDataset  # noqa: F401,E402
