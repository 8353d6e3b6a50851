"""wav2vec-S streaming encoder forward, B200-native (sm_100a CUDA kernels behind a C ABI).

The directory name carries a hyphen; import it as ``wav2vec_s_b200`` (the loader module of that
name at the repository root registers this directory as the package).
"""
from . import cabi
from .model import (BlockWiseWav2Vec2Model, OnlineW2V2TransformerEncoder, Wav2VecSModel, base_architecture,
                    sinusoidal_table)

__all__ = ["cabi", "Wav2VecSModel", "BlockWiseWav2Vec2Model", "OnlineW2V2TransformerEncoder",
           "base_architecture", "sinusoidal_table"]
