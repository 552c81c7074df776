"""Mirror of the part of the reference's ``utils`` package that lies on the engine's path: the
energy / momentum conservation metrics.  ``from utils.metrics import compute_energy_error`` (reference
scripts/evaluate.py:151) resolves here when this directory's parent is on ``sys.path`` in place of the reference's
``src``.  Plotting and the GNN error metrics of the reference's package are out of scope."""
