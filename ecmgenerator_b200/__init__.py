"""ecmgenerator_b200 - B200-native per-tick agent update of the ECMGenerator crowd engine.

Product code only: `host` (host-side world/planner helpers, libecmhost.so), `gpu` (ctypes binding of
the CUDA C ABI, libecmgpu.so), `scenarios` (synthetic worlds/crowds of BASELINE.json's configs).
The CPU oracle lives under /oracle and is test infrastructure; nothing here imports it.
"""
