"""Drop-in for the reference's params.py (star-imported by model.py; values bound at construction time).

Same names and values as the reference (params.py:5-50); only the VGGish rows matter to the B200 path, the rest
are kept so `from params import *` callers keep working.
"""

# ---- dataset (params.py:5-16)
DATASET_PATH = "UrbanSound8K/audio/"
METADATA_PATH = "UrbanSound8K/metadata/UrbanSound8K.csv"
MAX_SECONDS = 4
SR_VGGISH, SR_RESNET = 16_000, 22_050
SAMPLES_NUM_VGGISH = SR_VGGISH * MAX_SECONDS
SAMPLES_NUM_RESNET = SR_RESNET * MAX_SECONDS
TARGET_NAMES = [
    "air_conditioner", "car_horn", "children_playing", "dog_bark", "drilling",
    "engine_idling", "gun_shot", "jackhammer", "siren", "street_music",
]

# ---- model (params.py:22-32)
S_RESNET_SHAPE, S_VGGISH_SHAPE = (224, 224), (96, 64)  # CNN input image sizes
T = 10                    # bottleneck features (time steps) per clip
M_RESNET = 2048           # ResNet50 bottleneck width
M_VGGISH = 128            # VGGish embedding width
M_VGGISH_JB = 512 * 6 * 4  # VGGish conv features when just_bottlenecks=True
H = 600                   # hidden width of the attention head
DR = 0.4                  # dropout rate
K = 10                    # classes (UrbanSound8K); set model.K = 527 before building the head for AudioSet

# ---- training settings (params.py:39-50)
BATCH_SIZE = 8
NUM_EPOCHS = 25
FEATURE_EXTRACT = True
LR = 0.001
