import tensorflow as tf


def sanitize_seed(seed, salt=None, name=None):
    return seed


def split_seed(seed, n=2, salt=None, name=None):
    return [seed] * n


def uniform(shape, minval=0, maxval=None, dtype=tf.float32, seed=None, name=None):
    return tf.random.uniform(shape, dtype=dtype)
