"""Mirror of the reference package path ``global_recon`` (registry + config) for drop-in imports."""
