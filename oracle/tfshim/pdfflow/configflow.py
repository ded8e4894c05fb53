"""pdfflow.configflow stand-in: the dtype helpers madflow/config.py re-exports
(reference: python_package/madflow/config.py:37-47).  [EXT] pdfflow defines
float_me/int_me as tf.cast(x, DTYPE/DTYPEINT); the cast goes through
tf.convert_to_tensor, hence through float32 for bare Python floats."""
import os
import tensorflow as tf

DTYPE = tf.float64 if os.environ.get("PDFFLOW_FLOAT", "64") == "64" else tf.float32
DTYPEINT = tf.int32 if os.environ.get("PDFFLOW_INT", "32") == "32" else tf.int64


def run_eager(flag=True):
    return None


def float_me(x):
    return tf.cast(x, DTYPE)


def int_me(x):
    return tf.cast(x, DTYPEINT)


izero = int_me(0)
ione = int_me(1)
fzero = float_me(0.0)
fone = float_me(1.0)
