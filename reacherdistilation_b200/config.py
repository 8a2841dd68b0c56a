"""Constants of the reference's config module (/root/reference src/distilation/config.py:15-32), same names and values,
without its import-time side effects (timestamped mkdir, baselines imports).  Extra knobs for the batched path follow."""
import os

EPISODE_STEPS = 50
OBSPACE_SHAPE = 11
ACSPACE_SHAPE = 2
PDFLAT_SHAPE = 4
GAMMA = 0.99

TOTAL_EPISODES = 8000
STEPS_UNROLLED = 10
LSTM_BATCH_SIZE = 20
MLP_BATCH_SIZE = 20
NUM_UNITS = 200
KEEP_PROB = 0.5
MAX_CAPACITY = 10
TRAINING_EPOCHS = 1

# batched-path defaults (not in the reference: it runs one env, batch 1)
NUM_ENVS = int(os.environ.get("REACHER_B200_NUM_ENVS", 4096))
SEED = 0
# teacher logstd recorded in the reference fixture (tests/data/dataset.json, 't'[2:4])
TEACHER_LOGSTD = (-3.2939295768737793, -3.3629262447357178)

base_path = os.environ.get("REACHER_B200_DATA", os.path.join(os.path.expanduser("~"), "reacher", "data"))
