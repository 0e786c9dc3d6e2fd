"""Device-side caches of fitted objects (index handles, projector handles, flattened forests).

They live OUTSIDE the estimators' ``__dict__`` - keyed weakly by the owning object - so that
``transform`` / ``predict`` never change the estimator's attributes (scikit-learn's
``check_dict_unchanged``), ``clone`` / ``pickle`` never see them, and a handle dies with its owner.
"""

from __future__ import annotations

import weakref

_CACHE: "weakref.WeakKeyDictionary[object, dict]" = weakref.WeakKeyDictionary()


def get(owner, key, default=None):
    return _CACHE.get(owner, {}).get(key, default)


def put(owner, key, value):
    _CACHE.setdefault(owner, {})[key] = value
    return value


def drop(owner, *keys):
    slot = _CACHE.get(owner)
    if slot is None:
        return
    if not keys:
        slot.clear()
    for k in keys:
        slot.pop(k, None)
