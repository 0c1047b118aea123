"""wenet_b200 -- the Wenet receive hot path (fsk_demod | drs232_ldpc / wenet_ldpc) on B200: see DESIGN.md."""


def __getattr__(name):          # lazy: importing the package must not need numpy / the CUDA library
    if name == "Engine":
        from .engine import Engine
        return Engine
    if name == "MultiEngine":
        from .multi import MultiEngine
        return MultiEngine
    raise AttributeError(name)
