"""Mirror of the reference's `src` package (hot-path subset)."""
