"""audian_b200 -- B200-native derived-trace DSP path of bendalab/audian.

Spectrogram, Butterworth SOS filter, envelope and full-trace min/max as
hand-written sm_100a CUDA kernels behind a C ABI (include/audian_b200.h),
wrapped in drop-in replacements of audian's BufferedData traces.
"""

__version__ = '0.1.0'

from .buffereddata import BufferedData                    # noqa: F401
from .bufferedfilter import BufferedFilter                # noqa: F401
from .bufferedenvelope import BufferedEnvelope            # noqa: F401
from .bufferedspectrogram import BufferedSpectrogram      # noqa: F401
from .compresseddata import CompressedData                # noqa: F401
from .plugin import audian_b200_traces                    # noqa: F401
