"""The hydra-zen registration at the bottom of the plug-in modules: with an importable
``mmlearn.conf.external_store`` the fused loss and the fused retrieval metric register themselves
in the reference's groups (contrastive.py:19 ``modules/losses``, retrieval_recall.py:20
``modules/metrics``).  mmlearn itself is not installable here, so a stand-in store is injected."""
import importlib
import sys
import types


def test_external_store_registration(monkeypatch):
    calls = []

    def external_store(obj=None, **kw):
        calls.append((getattr(obj, "__name__", None), kw))
        return obj

    pkg = types.ModuleType("mmlearn")
    pkg.__path__ = []
    conf = types.ModuleType("mmlearn.conf")
    conf.external_store = external_store
    monkeypatch.setitem(sys.modules, "mmlearn", pkg)
    monkeypatch.setitem(sys.modules, "mmlearn.conf", conf)
    import mmlearn_b200.loss as loss_mod
    import mmlearn_b200.retrieval as retr_mod

    try:
        importlib.reload(loss_mod)
        importlib.reload(retr_mod)
        got = {name: kw for name, kw in calls}
        assert got["ContrastiveLoss"]["name"] == "FusedContrastiveLoss"
        assert got["ContrastiveLoss"]["group"] == "modules/losses"
        assert got["RetrievalRecallAtK"]["name"] == "FusedRetrievalRecallAtK"
        assert got["RetrievalRecallAtK"]["group"] == "modules/metrics"
    finally:
        monkeypatch.delitem(sys.modules, "mmlearn")
        monkeypatch.delitem(sys.modules, "mmlearn.conf")
        importlib.reload(loss_mod)
        importlib.reload(retr_mod)
        import mmlearn_b200

        importlib.reload(mmlearn_b200)
