"""Load the reference's own `app/main.py` with its absent services stubbed.

TEST INFRASTRUCTURE ONLY.  Works only where /root/reference exists (the build
container); nothing that runs on the GPU box imports this.  It is used by
`oracle/make_golden.py` to produce the committed fixtures under tests/golden/
and by the CPU tests (skipped when the reference is absent) to compare the
oracle with the reference live.

Stub surface: `redis.Redis` as an in-memory list,
`opensearchpy.OpenSearch` as a fake whose `search` does exact numpy cosine +
stable sort over what `bulk` stored, empty `spacy`, `langchain.memory`.
After loading, `cosine_similarity`, `lfu_cache_get`, `lfu_cache_put`,
`_remove_least_frequent_item`, `OpenSearchIndexer` are the reference's code
objects, unmodified.
"""
from __future__ import annotations

import contextlib
import importlib.util
import io
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("SQE_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "app", "main.py"))


class FakeRedis:
    """The five list commands the reference calls (main.py:69,95,117,125,128)."""

    def __init__(self, *a, **kw):
        self.lists = {}

    def _l(self, name):
        return self.lists.setdefault(name, [])

    def lrange(self, name, start, stop):
        l = self._l(name)
        stop = len(l) if stop == -1 else stop + 1
        return list(l[start:stop])

    def lset(self, name, index, value):
        self._l(name)[index] = value

    def llen(self, name):
        return len(self._l(name))

    def lpush(self, name, value):
        self._l(name).insert(0, value)

    def lrem(self, name, count, value):
        l = self._l(name)
        if value in l:
            l.remove(value)
            return 1
        return 0


class FakeOpenSearch:
    """Stores what `bulk` sends; `search` = exact cosine on the stored
    (already normalised) vectors, stable sort, best first."""

    def __init__(self, **kw):
        self.docs = []          # list of (_id, _source)
        self.fail_info = kw.pop("_fail_info", False)
        self.last_query = None
        outer = self

        class _Indices:
            def exists(self, *a, **k):
                return True

            def create(self, *a, **k):
                return {}

        self.indices = _Indices()

    def info(self):
        if self.fail_info:
            raise RuntimeError("stub: no OpenSearch in this container")
        return {"version": "stub"}

    def count(self, index=None):
        return {"count": len(self.docs)}

    def search(self, index=None, body=None):
        self.last_query = body
        k = body["size"]
        q = np.asarray(body["query"]["knn"]["embedding"]["vector"], dtype=np.float32)
        if not self.docs:
            return {"hits": {"hits": []}}
        d = np.asarray([s["embedding"] for _, s in self.docs], dtype=np.float32)
        scores = d @ q
        order = np.argsort(-scores, kind="stable")[:k]
        return {"hits": {"hits": [
            {"_id": self.docs[i][0], "_score": float(scores[i]), "_source": self.docs[i][1]}
            for i in order]}}


def _fake_bulk(client, actions):
    for a in actions:
        # an "index" action REPLACES the document that already has this _id (OpenSearch bulk API);
        # the replaced document keeps its position, so tie order does not change
        for j, (known, _) in enumerate(client.docs):
            if known == a["_id"]:
                client.docs[j] = (a["_id"], a["_source"])
                break
        else:
            client.docs.append((a["_id"], a["_source"]))
    return len(actions), []


def load_reference_main(quiet: bool = True):
    """Return the reference `main` module (fresh copy each call)."""
    if not reference_available():
        raise FileNotFoundError(REFERENCE_ROOT)

    def mod(name, **kw):
        m = types.ModuleType(name)
        m.__dict__.update(kw)
        sys.modules[name] = m
        return m

    saved = {k: sys.modules.get(k) for k in
             ("redis", "opensearchpy", "opensearchpy.helpers", "spacy", "langchain",
              "langchain.memory")}
    mod("redis", Redis=FakeRedis)
    osp = mod("opensearchpy", OpenSearch=lambda **kw: FakeOpenSearch(_fail_info=True, **kw),
              RequestsHttpConnection=object)
    osp.helpers = mod("opensearchpy.helpers", bulk=_fake_bulk)
    mod("spacy")
    lc = mod("langchain")
    lc.memory = mod("langchain.memory", ConversationBufferMemory=object)
    try:
        spec = importlib.util.spec_from_file_location(
            "sqe_reference_main", os.path.join(REFERENCE_ROOT, "app", "main.py"))
        m = importlib.util.module_from_spec(spec)
        with contextlib.redirect_stdout(io.StringIO() if quiet else sys.stdout):
            spec.loader.exec_module(m)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return m


class FakeMultiIndexOpenSearch:
    """Per-index document store for the upload micro-service (embedding_gen.py): `bulk` actions
    carry `_index`; an "index" action replaces the document with the same `_id` in place."""

    def __init__(self, **kw):
        self.by_index = {}
        outer = self

        class _Indices:
            def exists(self, index=None, **k):
                return index in outer.by_index

            def create(self, index=None, body=None, **k):
                outer.by_index.setdefault(index, [])
                outer.bodies = getattr(outer, "bodies", {})
                outer.bodies[index] = body
                return {}

        self.indices = _Indices()


def _fake_multi_bulk(client, actions):
    for a in actions:
        docs = client.by_index.setdefault(a["_index"], [])
        for j, (known, _) in enumerate(docs):
            if known == a["_id"]:
                docs[j] = (a["_id"], a["_source"])
                break
        else:
            docs.append((a["_id"], a["_source"]))
    return len(actions), []


def load_reference_embedding_gen(base_index_name: str = "docs", quiet: bool = True):
    """Return the reference's `app/embedding_gen.py` (the upload micro-service) with Postgres and
    OpenSearch stubbed; `os_client` is a FakeMultiIndexOpenSearch.  Loaded from a scratch working
    directory because the module creates `uploads/` relative to the cwd at import."""
    import tempfile
    if not reference_available():
        raise FileNotFoundError(REFERENCE_ROOT)

    def mod(name, **kw):
        m = types.ModuleType(name)
        m.__dict__.update(kw)
        sys.modules[name] = m
        return m

    names = ("opensearchpy", "opensearchpy.helpers", "psycopg2", "psycopg2.extras", "asyncpg")
    saved = {k: sys.modules.get(k) for k in names}
    osp = mod("opensearchpy", OpenSearch=lambda **kw: FakeMultiIndexOpenSearch(**kw), RequestsHttpConnection=object)
    osp.helpers = mod("opensearchpy.helpers", bulk=_fake_multi_bulk)
    pg = mod("psycopg2")
    pg.extras = mod("psycopg2.extras")
    mod("asyncpg")
    cwd = os.getcwd()
    old_env = os.environ.get("OPENSEARCH_INDEX_NAME")
    os.environ["OPENSEARCH_INDEX_NAME"] = base_index_name
    try:
        with tempfile.TemporaryDirectory() as tmp:
            os.chdir(tmp)
            spec = importlib.util.spec_from_file_location(
                "sqe_reference_embedding_gen", os.path.join(REFERENCE_ROOT, "app", "embedding_gen.py"))
            m = importlib.util.module_from_spec(spec)
            with contextlib.redirect_stdout(io.StringIO() if quiet else sys.stdout):
                spec.loader.exec_module(m)
            os.chdir(cwd)
    finally:
        os.chdir(cwd)
        if old_env is None:
            os.environ.pop("OPENSEARCH_INDEX_NAME", None)
        else:
            os.environ["OPENSEARCH_INDEX_NAME"] = old_env
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return m
