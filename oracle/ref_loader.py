"""TEST INFRASTRUCTURE ONLY -- loads the *unmodified* reference loss by file path.

mmlearn cannot be imported as a package in this image (hydra_zen, lightning, torchmetrics,
timm, peft are not installed and there is no wheel for them; hatchling is missing too, so
``pip install /root/reference`` fails).  The reference source files themselves are pure
Python on top of torch, so they run in place once the missing imports are stubbed:

* ``hydra_zen.store``            -> identity decorator (``@store`` and ``@store(group=...)``)
* ``torchmetrics.utilities.compute._safe_matmul`` -> restated from torchmetrics 1.6.2
  (pinned in the reference's ``uv.lock``): fp16 inputs are up-cast, everything else is
  ``x @ y.T``.  Call sites: ``mmlearn/modules/losses/contrastive.py:327,330,334,339,340``.
* ``lightning.fabric.utilities.rank_zero_warn`` -> ``warnings.warn``

Only available where ``/root/reference`` exists (the authoring container).  It is used by
``oracle/make_golden.py`` to produce ``tests/golden/*.npz`` and by the CPU tests that pin
``oracle/contrastive_oracle.py`` against the real reference.  Nothing in the product imports it.
"""

from __future__ import annotations

import importlib.util
import os
import sys
import types
import warnings
from dataclasses import dataclass

REFERENCE_ROOT = os.environ.get("MMLEARN_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(
        os.path.join(REFERENCE_ROOT, "mmlearn/modules/losses/contrastive.py")
    )


def _load(name: str, relpath: str) -> types.ModuleType:
    spec = importlib.util.spec_from_file_location(
        name, os.path.join(REFERENCE_ROOT, relpath)
    )
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


_CACHE: dict[str, types.ModuleType] = {}


def load_reference() -> dict[str, types.ModuleType]:
    """Return {"contrastive": module, "example": module, "modalities": module}."""
    if _CACHE:
        return _CACHE
    if not reference_available():
        raise RuntimeError(f"reference sources not found under {REFERENCE_ROOT}")
    import torch

    saved = {k: sys.modules.get(k) for k in list(sys.modules) if k.split(".")[0] in
             ("mmlearn", "hydra_zen", "torchmetrics", "lightning")}

    # --- stubs -------------------------------------------------------------------------
    hz = types.ModuleType("hydra_zen")

    def store(*args, **kwargs):
        if len(args) == 1 and callable(args[0]) and not kwargs:
            return args[0]
        return lambda obj: obj

    hz.store = store
    hz.MISSING = object()
    sys.modules["hydra_zen"] = hz

    tm = types.ModuleType("torchmetrics")
    tmu = types.ModuleType("torchmetrics.utilities")
    tmc = types.ModuleType("torchmetrics.utilities.compute")

    def _safe_matmul(x, y):
        if x.dtype == torch.float16 or y.dtype == torch.float16:
            return (x.float() @ y.T.float()).half()
        return x @ y.T

    tmc._safe_matmul = _safe_matmul
    tm.utilities = tmu
    tmu.compute = tmc
    sys.modules["torchmetrics"] = tm
    sys.modules["torchmetrics.utilities"] = tmu
    sys.modules["torchmetrics.utilities.compute"] = tmc

    li = types.ModuleType("lightning")
    lf = types.ModuleType("lightning.fabric")
    lfu = types.ModuleType("lightning.fabric.utilities")
    lfu.rank_zero_warn = lambda msg, *a, **k: warnings.warn(msg)
    li.fabric = lf
    lf.utilities = lfu
    sys.modules["lightning"] = li
    sys.modules["lightning.fabric"] = lf
    sys.modules["lightning.fabric.utilities"] = lfu

    # package placeholders so the heavy __init__ files are skipped
    for pkg in ("mmlearn", "mmlearn.datasets", "mmlearn.datasets.core", "mmlearn.tasks",
                "mmlearn.modules", "mmlearn.modules.losses"):
        m = types.ModuleType(pkg)
        m.__path__ = []  # type: ignore[attr-defined]
        sys.modules[pkg] = m

    modalities = _load("mmlearn.datasets.core.modalities", "mmlearn/datasets/core/modalities.py")
    example = _load("mmlearn.datasets.core.example", "mmlearn/datasets/core/example.py")
    core = sys.modules["mmlearn.datasets.core"]
    core.find_matching_indices = example.find_matching_indices
    core.Modalities = modalities.Modalities
    core.modalities = modalities
    core.example = example

    # LossPairSpec lives in the task file, which drags in Lightning; the dataclass is 3 lines
    # (mmlearn/tasks/contrastive_pretraining.py:44-52) and is only used as a type + attribute bag.
    cp = types.ModuleType("mmlearn.tasks.contrastive_pretraining")

    @dataclass
    class LossPairSpec:
        modalities: tuple
        weight: float = 1.0

    cp.LossPairSpec = LossPairSpec
    sys.modules["mmlearn.tasks.contrastive_pretraining"] = cp

    contrastive = _load(
        "mmlearn.modules.losses.contrastive", "mmlearn/modules/losses/contrastive.py"
    )
    _CACHE.update(
        contrastive=contrastive, example=example, modalities=modalities, task_stub=cp
    )
    return _CACHE


def load_reference_retrieval() -> types.ModuleType:
    """The unmodified ``mmlearn/modules/metrics/retrieval_recall.py`` under import stubs.

    torchmetrics is not installable here, so the pieces the file imports are restated from
    torchmetrics 1.6.2 (pinned in the reference's ``uv.lock``): a minimal ``Metric`` base
    (``add_state``, ``reset``, single-process ``distributed_available_fn``), ``dim_zero_cat``,
    ``_check_same_shape``, ``_retrieval_aggregate``, ``_safe_matmul``.  Everything that decides the
    metric's value -- ``update``, ``compute``, ``_process_batch``, ``_recall_at_k`` -- is the
    reference's own code."""
    if "retrieval" in _CACHE:
        return _CACHE["retrieval"]
    load_reference()
    import torch

    tm = sys.modules["torchmetrics"]

    class Metric(torch.nn.Module):
        def __init__(self, **kwargs):
            super().__init__()
            self._defaults = {}
            self.distributed_available_fn = lambda: False
            self.process_group = None

        def add_state(self, name, default, dist_reduce_fx=None):
            self._defaults[name] = default
            setattr(self, name, [] if isinstance(default, list) else default.clone())

        def reset(self):
            for name, default in self._defaults.items():
                setattr(self, name, [] if isinstance(default, list) else default.clone())

    tm.Metric = Metric

    def mod(name):
        m = types.ModuleType(name)
        sys.modules[name] = m
        return m

    base = mod("torchmetrics.retrieval.base")
    mod("torchmetrics.retrieval").base = base

    def _retrieval_aggregate(values, aggregation="mean", dim=None):
        if aggregation == "mean":
            return values.mean() if dim is None else values.mean(dim=dim)
        if aggregation == "median":
            return values.median() if dim is None else values.median(dim=dim).values
        if aggregation == "min":
            return values.min() if dim is None else values.min(dim=dim).values
        if aggregation == "max":
            return values.max() if dim is None else values.max(dim=dim).values
        return aggregation(values, dim=dim)

    base._retrieval_aggregate = _retrieval_aggregate
    checks = mod("torchmetrics.utilities.checks")

    def _check_same_shape(preds, target):
        if preds.shape != target.shape:
            raise RuntimeError(
                "Predictions and targets are expected to have the same shape, "
                f"but got {preds.shape} and {target.shape}.")

    checks._check_same_shape = _check_same_shape
    data = mod("torchmetrics.utilities.data")

    def dim_zero_cat(x):
        if isinstance(x, torch.Tensor):
            return x
        x = [y.unsqueeze(0) if y.numel() == 1 and y.ndim == 0 else y for y in x]
        if not x:
            raise ValueError("No samples to concatenate")
        return torch.cat(x, dim=0)

    data.dim_zero_cat = dim_zero_cat
    distm = mod("torchmetrics.utilities.distributed")
    distm.gather_all_tensors = lambda t, group=None: [t]
    sys.modules["mmlearn.modules.metrics"] = types.ModuleType("mmlearn.modules.metrics")
    _CACHE["retrieval"] = _load("mmlearn.modules.metrics.retrieval_recall",
                                "mmlearn/modules/metrics/retrieval_recall.py")
    return _CACHE["retrieval"]


def load_reference_zero_shot() -> types.ModuleType:
    """The unmodified ``mmlearn/tasks/zero_shot_classification.py`` under import stubs.

    What the file imports but this image lacks is stubbed: ``lightning.pytorch`` (type names and
    ``move_data_to_device``), the torchmetrics class names (``Accuracy`` ... -- never
    instantiated by the fixture generator, which injects a capturing collection) and
    ``CombinedDataset``.  ``evaluation_step`` -- the scoring arithmetic the fused path replaces --
    is the reference's own code."""
    if "zero_shot" in _CACHE:
        return _CACHE["zero_shot"]
    load_reference()

    def mod(name):
        m = sys.modules.get(name) or types.ModuleType(name)
        sys.modules[name] = m
        return m

    lp = mod("lightning.pytorch")
    lp.LightningModule = object
    sys.modules["lightning"].pytorch = lp
    lpu = mod("lightning.pytorch.utilities")
    lpu.move_data_to_device = lambda data, device: data
    lpr = mod("lightning.pytorch.utilities.rank_zero")
    lpr.rank_zero_warn = lambda msg, *a, **k: warnings.warn(msg)
    tm = sys.modules["torchmetrics"]
    for name in ("AUROC", "Accuracy", "F1Score", "MetricCollection", "Precision", "Recall"):
        setattr(tm, name, type(name, (), {}))
    core = sys.modules["mmlearn.datasets.core"]
    core.CombinedDataset = type("CombinedDataset", (), {})
    _load("mmlearn.tasks.hooks", "mmlearn/tasks/hooks.py")
    _CACHE["zero_shot"] = _load("mmlearn.tasks.zero_shot_classification",
                                "mmlearn/tasks/zero_shot_classification.py")
    return _CACHE["zero_shot"]
