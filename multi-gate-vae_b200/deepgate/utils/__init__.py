from . import dag_utils
from . import utils
