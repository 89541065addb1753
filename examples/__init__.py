"""The reference's example / test programs written in the xgrid DSL plus synthetic-input builders:
inputs of the tests, ``bench.py`` and ``__graft_entry__.smoke()`` -- not part of the product package."""
