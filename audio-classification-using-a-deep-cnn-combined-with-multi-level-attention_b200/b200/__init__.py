"""B200-native host layer: ctypes binding (`_lib`), tensor-level wrappers (`engine`), synthetic data (`synth`),
batch sharding across GPUs (`sharding`) and the in-tree build (`build`)."""
from ._lib import B200Error, LIB_PATH  # noqa: F401
