"""vidchapters_b200 — B200-native Vid2Seq train step (drop-in for the reference's model.vid2seq.Vid2Seq)."""
from .config import CONFIGS, T5_BASE, T5_LARGE, TINY, TINY_PROJ  # noqa: F401
from .graphed import GraphedTrainStep  # noqa: F401
from .optim import Vid2SeqAdam  # noqa: F401
from .vid2seq import Vid2Seq, _get_tokenizer, build_vid2seq_model  # noqa: F401
from .data import DenseVideoCaptioningDataset, PinnedBatcher, YTDataset, collate_dvc  # noqa: F401  (host input pipeline, §8f N3)
