"""Drop-in mirror of the reference's `modules` package for the forward hot path: same module file names, class
names, constructor arguments, state_dict keys and forward signatures; the forwards run on
libmvlt_b200.so."""
