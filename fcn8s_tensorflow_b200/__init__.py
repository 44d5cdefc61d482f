"""fcn8s_tensorflow_b200 -- B200-native FCN-8s forward+backward engine behind the FCN8s class surface of
pierluigiferrari/fcn8s_tensorflow."""
__version__ = "0.1.0"
