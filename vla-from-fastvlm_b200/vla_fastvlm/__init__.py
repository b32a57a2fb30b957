"""B200-native FastVLA policy forward behind the VLA-from-FastVLM plugin surface.

`vla_fastvlm.fastvla`, `vla_fastvlm.model` and (with LeRobot installed) `vla_fastvlm.lerobot_fastvla`
keep the reference's module, class and function names; the arithmetic runs in libfvla (sm_100a)."""

__all__ = ["FastVLAConfig", "FastVLAPolicy"]


def __getattr__(name):  # lazy: importing the package must not require a GPU or the built library
    if name in __all__:
        from . import fastvla

        return getattr(fastvla, name)
    raise AttributeError(name)
