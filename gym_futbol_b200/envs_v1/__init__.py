# v1 (`Futbol`, pymunk physics) drop-in: lands with the v1 kernels.
