"""Stand-in for the CUDA engine backed by the CPU oracle, plus a pytest plugin that installs it for a
whole session (`-p oracle.engine_standin`).  TEST INFRASTRUCTURE ONLY: CPU tests use it to exercise the
host pipeline (parsing, dedup, cache re-indexing, labelling, output) without a GPU, and
tests/test_reference_suite.py uses the plugin to run the reference's own test-suite against this
package on a GPU-less box.  Never reachable from the product."""
from __future__ import annotations

import numpy as np


class OracleEngine:
    """Stand-in for breakfast_b200.engine.components_* backed by the CPU oracle.  CPU tests use it to
    exercise the host pipeline (parsing, dedup, cache re-indexing, labelling, output) without a GPU.
    Test infrastructure only: installed with monkeypatch, never reachable from the product."""

    @staticmethod
    def full(indptr, indices, n_cols, max_dist, want_edges=False, device=None, engine=None):
        import oracle
        from breakfast_b200.engine import ClusterResult
        src, dst = oracle.edges(indptr, indices, max_dist)
        labels = oracle.components(len(indptr) - 1, src, dst)
        return ClusterResult(labels, {}, (src, dst) if want_edges else None)

    @staticmethod
    def incremental(indptr, indices, n_cols, max_dist, new_rows, list_indptr, list_members, want_edges=False,
                    device=None, engine=None):
        import oracle
        from breakfast_b200.engine import ClusterResult
        new_rows = np.unique(np.asarray(new_rows, dtype=np.int32))
        n = len(indptr) - 1
        if new_rows.size:
            src, dst = oracle.edges(indptr, indices, max_dist, queries=new_rows)
        else:
            src = dst = np.zeros(0, np.int32)
        labels = oracle.components(n, src, dst, list_indptr, list_members)
        return ClusterResult(labels, {}, (src, dst) if want_edges else None)

    @classmethod
    def install(cls, monkeypatch):
        from breakfast_b200 import engine
        monkeypatch.setattr(engine, "components_full", cls.full)
        monkeypatch.setattr(engine, "components_incremental", cls.incremental)


class HashJoinEngine:
    """Like OracleEngine.full, but backed by the second oracle (oracle/hashjoin.py, max_dist <= 2): fast enough for
    the scale goldens (tens of thousands of profiles) on the CPU."""

    @staticmethod
    def full(indptr, indices, n_cols, max_dist, want_edges=False, device=None, engine=None):
        import oracle
        from oracle import hashjoin
        from breakfast_b200.engine import ClusterResult
        src, dst = hashjoin.edges(indptr, indices, max_dist)
        labels = oracle.components(len(indptr) - 1, src, dst)
        return ClusterResult(labels, {}, (src, dst) if want_edges else None)

    @classmethod
    def install(cls, monkeypatch):
        from breakfast_b200 import engine
        monkeypatch.setattr(engine, "components_full", cls.full)


def pytest_configure(config):
    from breakfast_b200 import engine
    engine.components_full = OracleEngine.full
    engine.components_incremental = OracleEngine.incremental
