"""Mirror of the reference's `src/model` package for the render hot path."""
