"""Same names/values as the reference's ``dataset/waveform/waveform_configs.py``."""
from ..common_config import *  # noqa: F401,F403
from ..common_config import frame_size, working_sample_rate, hop_size, audio_channels
from ...utils.common import human_format

cfg_descriptor = f"WaveForm_SaR-{human_format(working_sample_rate)}_FrS-{human_format(frame_size)}" \
                 f"_HoS-{human_format(hop_size)}_Ch-{audio_channels}"
