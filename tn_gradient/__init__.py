"""Drop-in ``tn_gradient`` namespace backed by sow_b200 (same module paths as antoine311200/sow)."""
