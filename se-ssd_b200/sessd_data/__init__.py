"""sessd_data -- seeded synthetic inputs, seeded model parameters and layer tables of the SE-SSD car model.

Pure numpy / torch-CPU: importing this package does NOT load libsessd_b200.so, so the CPU reference arm of bench.py and the
oracle-side tests can share exactly the same inputs and weights as the CUDA path without touching the product library.
"""
