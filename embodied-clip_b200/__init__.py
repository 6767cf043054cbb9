"""embclip_b200: B200-native (sm_100a) implementation of EmbCLIP's data-parallel hot path --
the frozen CLIP visual encoder behind AllenAct's Preprocessor surface, computed by hand-written
tcgen05 / TMA CUDA kernels in ``libembclip_b200.so`` (C ABI: include/embclip_b200.h).

There is no CPU or PyTorch fallback: importing the compute classes without the built library
raises, and so does running them without a CUDA device."""
__version__ = "0.1.0"
