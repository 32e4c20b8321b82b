"""Empty stand-in for the top-level `import tensorflow` at tfmesos/server.py:7.
Only the reference's cmd (replica) mode is exercised with it; the
tf.train.Server branch (server.py:51-66) is the part this repo replaces."""
