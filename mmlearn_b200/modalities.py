"""Interface-only mirrors of the two reference types the loss touches.

* ``LossPairSpec``  -- mmlearn/tasks/contrastive_pretraining.py:44-52
* ``Modalities.get_modality(name).name/.embedding`` -- mmlearn/datasets/core/modalities.py:15-82,
  227-240 (``rgb`` -> ``rgb_embedding``; names are lower-cased; unknown names raise KeyError).

When mmlearn itself is importable its own objects are used, so custom modalities registered
there are honoured; otherwise this small lookalike is used.
"""

from __future__ import annotations

from dataclasses import dataclass

try:  # pragma: no cover - mmlearn is not installable in the build image
    from mmlearn.datasets.core.modalities import Modalities  # type: ignore
    from mmlearn.tasks.contrastive_pretraining import LossPairSpec  # type: ignore
except Exception:  # noqa: BLE001

    @dataclass
    class LossPairSpec:  # type: ignore[no-redef]
        """Specification for a pair of modalities to compute the contrastive loss."""

        modalities: tuple
        weight: float = 1.0

    @dataclass
    class Modality:
        name: str

        def __post_init__(self) -> None:
            self.name = self.name.lower()

        @property
        def embedding(self) -> str:
            return f"{self.name}_embedding"

    class _Registry:
        """Name -> Modality; the reference pre-registers rgb/depth/thermal/text/audio/video
        (modalities.py:10)."""

        def __init__(self) -> None:
            self._m: dict[str, Modality] = {}
            for n in ("rgb", "depth", "thermal", "text", "audio", "video"):
                self.register_modality(n)

        def register_modality(self, name: str, modality_specific_properties=None) -> None:
            self._m[name.lower()] = Modality(name)

        def get_modality(self, name: str) -> Modality:
            try:
                return self._m[name.lower()]
            except KeyError:
                raise KeyError(f"Modality '{name.lower()}' not found.") from None

        def has_modality(self, name: str) -> bool:
            return name.lower() in self._m

        def list_modalities(self):
            return list(self._m.values())

    Modalities = _Registry()  # type: ignore[assignment]

__all__ = ["LossPairSpec", "Modalities"]
