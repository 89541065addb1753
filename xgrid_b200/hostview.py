"""NumPy views that tell their ``Grid`` before they are written.

In the reference ``Grid.boundary`` and ``Grid.now`` are plain NumPy arrays that
ARE the kernel's storage (xgrid/xgrid/__init__.py:38-41,70-72): a program may
keep one (``b = g.boundary``) and write through it between kernel calls.  Here
the kernel's storage is in HBM, so the arrays handed out are ``ndarray``
subclass views whose mutating entry points -- item assignment, in-place
operators / ``out=`` ufuncs, ``fill`` / ``put`` and the ``np.copyto`` family --
first call a hook on the owning grid:

* a mask view marks the mask as touched (the next kernel call re-compares and,
  if it changed, re-compiles it) and runs any deferred kernel calls first, so
  queued steps still see the mask they were called with;
* a level view makes the host copy the truth again (downloading the level
  first if the device holds newer data), so the next kernel call uploads it.

Reads through a retained view are NOT refreshed: read ``grid.now`` again after
kernel calls.  Everything computed from a view is a plain ``ndarray``.
"""
from __future__ import annotations

import weakref

import numpy as np

_WRITERS = {"copyto", "put", "place", "putmask", "put_along_axis", "fill_diagonal"}


def _plain(x):
    return x.view(np.ndarray) if isinstance(x, HostView) else x


class HostView(np.ndarray):
    _hook = None                     # weak bound method of the owner, called BEFORE a write

    def __array_finalize__(self, obj) -> None:
        self._hook = getattr(obj, "_hook", None)

    def _touch(self) -> None:
        hook = self._hook
        if hook is not None:
            fn = hook()
            if fn is not None:
                fn(self)

    def __setitem__(self, key, value) -> None:
        self._touch()
        np.ndarray.__setitem__(self.view(np.ndarray), key, _plain(value))

    def fill(self, value) -> None:
        self._touch()
        self.view(np.ndarray).fill(value)

    def put(self, *args, **kwargs) -> None:
        self._touch()
        self.view(np.ndarray).put(*args, **kwargs)

    def sort(self, *args, **kwargs) -> None:
        self._touch()
        self.view(np.ndarray).sort(*args, **kwargs)

    def __array_ufunc__(self, ufunc, method, *inputs, out=None, **kwargs):
        if out is not None:
            for o in out:
                if isinstance(o, HostView):
                    o._touch()
            kwargs["out"] = tuple(_plain(o) for o in out)
        if method == "at" and inputs and isinstance(inputs[0], HostView):
            inputs[0]._touch()
        result = getattr(ufunc, method)(*[_plain(i) for i in inputs], **kwargs)
        if out is not None and len(out) == 1 and isinstance(out[0], HostView):
            return out[0]            # `b += 1` must rebind the name to the same view
        return result

    def __array_function__(self, func, types, args, kwargs):
        if func.__name__ in _WRITERS and args and isinstance(args[0], HostView):
            args[0]._touch()
        return super().__array_function__(func, types, args, kwargs)

    def __reduce__(self):
        return self.view(np.ndarray).__reduce__()


def make(array: np.ndarray, hook) -> HostView:
    """View of `array` that calls the bound method `hook(view)` (held weakly) before every write."""
    v = array.view(HostView)
    v._hook = weakref.WeakMethod(hook)
    return v
