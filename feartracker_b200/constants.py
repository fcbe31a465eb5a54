"""Dictionary keys of the network outputs (same string values as the reference's
model_training/utils/constants.py:1,3 so callers can index either way)."""
TARGET_CLASSIFICATION_KEY = "TARGET_CLASSIFICATION_KEY"
TARGET_REGRESSION_LABEL_KEY = "TARGET_REGRESSION_LABEL_KEY"
