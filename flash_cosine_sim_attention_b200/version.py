# same version and extension-module name as the reference release this package drops in for
# (reference version.py:1-3): the torch extension built from csrc/torch_ext.cpp carries this name
__version__ = '0.1.40'

__cuda_pkg_name__ = f'flash_cosine_sim_attention_cuda_{__version__.replace(".", "_")}'

# the C-ABI library the extension module binds
__abi_library__ = 'libfcsa_b200'
