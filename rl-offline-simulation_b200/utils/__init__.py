from .dataset_utils import record_dataset_in_memory, load_h5_dataset, get_keys
from .prob_utils import get_uniform_dist, sample_dist
