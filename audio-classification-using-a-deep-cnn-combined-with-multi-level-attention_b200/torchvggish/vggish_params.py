"""Drop-in for torchvggish/vggish_params.py: the VGGish constants, same names and values
(reference vggish_params.py:22-53).  The CUDA front end hard-codes the derived numbers (400-sample window,
160-sample hop, 512-point DFT, 64 mel bands 125-7500 Hz, log offset 0.01, 96-frame examples) and
tests/test_abi_host.py::test_constants_match_reference_values checks this file against the reference's values and
tests/test_abi_host.py::test_frame_arithmetic_matches_reference the derived frame counts.
"""

# architecture
NUM_FRAMES, NUM_BANDS, EMBEDDING_SIZE = 96, 64, 128

# feature / example generation
SAMPLE_RATE = 16000
STFT_WINDOW_LENGTH_SECONDS, STFT_HOP_LENGTH_SECONDS = 0.025, 0.010
NUM_MEL_BINS = NUM_BANDS
MEL_MIN_HZ, MEL_MAX_HZ = 125, 7500
LOG_OFFSET = 0.01
EXAMPLE_WINDOW_SECONDS = EXAMPLE_HOP_SECONDS = 0.96   # 96 frames of 10 ms, no overlap

# embedding post-processing
PCA_EIGEN_VECTORS_NAME, PCA_MEANS_NAME = "pca_eigen_vectors", "pca_means"
QUANTIZE_MIN_VAL, QUANTIZE_MAX_VAL = -2.0, +2.0

# training hyper-parameters of the original TF release (unused by the forward path)
INIT_STDDEV, LEARNING_RATE, ADAM_EPSILON = 0.01, 1e-4, 1e-8

# TF graph names (kept for importers)
INPUT_OP_NAME = "vggish/input_features"
INPUT_TENSOR_NAME = INPUT_OP_NAME + ":0"
OUTPUT_OP_NAME = "vggish/embedding"
OUTPUT_TENSOR_NAME = OUTPUT_OP_NAME + ":0"
AUDIO_EMBEDDING_FEATURE_NAME = "audio_embedding"
