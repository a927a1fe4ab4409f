"""Same names/values as the reference's ``dataset/spectogram/spectogram_configs.py:5-14``."""
import numpy as np
from ..common_config import *  # noqa: F401,F403
from ..common_config import frame_size, working_sample_rate, frames_per_second, hop_size, audio_channels
from ...utils.common import human_format

NFFT = 2 ** int(np.ceil(np.log2(frame_size)))
mel_bins = 64
mel_min_freq = 20
mel_max_freq = working_sample_rate // 2

train_crop_size = frames_per_second * 10

cfg_descriptor = f"Spectogram_SaR-{human_format(working_sample_rate)}_FrS-{human_format(frame_size)}" \
                 f"_HoS-{human_format(hop_size)}_Mel-{mel_bins}_Ch-{audio_channels}"
