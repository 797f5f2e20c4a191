"""B200-native statevector hot path for Qiskit Terra 0.18's BasicAer simulators.

    simulator.get_simulator(name).run(qobj_dict)      qiskit-free entry point (Qobj wire dict in, result dict out)
    provider.B200Aer.get_backend(name)                BackendV1 plugin (needs qiskit)
    engine.DeviceState                                device-resident state driving libqsv (include/qsv.h)
    statevector.DeviceStatevector                     Statevector-like handle on the device (Pauli expectation values,
                                                      probabilities, sampling without the 2^n-amplitude copy out)
    dist.distributed_simulator(name)                  the same simulators on a state sharded over the GPUs of a node

There is no CPU fallback: every path ends in the CUDA library ``libqsv.so`` built by ``__graft_entry__.build()``.
"""
__version__ = "0.1.0"
