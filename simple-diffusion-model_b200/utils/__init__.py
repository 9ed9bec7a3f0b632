"""Host-side helpers behind the reference names (progress bar, plots, checkpoints)."""
