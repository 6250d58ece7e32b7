"""auditory_b200 -- B200-native (sm_100a) implementation of emer/auditory's
speech-feature hot path: sound.SndEnv -> dft.Filter -> mel.FilterDft
[-> mel.CepstrumDct] -> agabor.Convolve, behind the C-ABI in
include/auditory_b200.h.  The modules mirror the reference's Go packages."""
from . import agabor, dft, kwta, mel, sound, synth  # noqa: F401
from ._lib import AudError, AudParams, lib  # noqa: F401
from .pipeline import Pipeline, process_host_multi  # noqa: F401
from .sound import SndEnv, Wave, MSecToSamples  # noqa: F401

__all__ = ["agabor", "dft", "kwta", "mel", "sound", "synth", "Pipeline", "process_host_multi", "SndEnv", "Wave", "MSecToSamples", "AudError", "AudParams", "lib"]
