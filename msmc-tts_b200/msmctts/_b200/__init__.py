"""B200 backend of the msmctts.networks hot path: ctypes loader (lib) + autograd glue (functional)."""
