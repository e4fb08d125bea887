"""Host side of the B200 step pipeline: model descriptors and the ctypes binding."""
