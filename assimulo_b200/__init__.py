"""assimulo_b200 -- B200-native ensemble integrator behind Assimulo's Explicit_Problem + solver.simulate()
surface.  The product is ``assimulo_b200.gpu`` (drop-in for an ``assimulo.gpu`` package)
over ``assimulo_b200/lib/libassimulo_cuda.so`` (C ABI: include/assimulo_cuda.h)."""
__version__ = "0.1.0"
