__version__ = '0.1.0'

# name of the C-ABI library this package binds (reference: version.py:3 named the pybind module)
__cuda_pkg_name__ = 'libfcsa_b200'
