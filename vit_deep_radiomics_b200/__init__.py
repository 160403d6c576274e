"""vit_deep_radiomics_b200 -- B200-native (sm_100a) implementation of the hot path of
larosi/vit-deep-radiomics: ViT dense-descriptor extraction -> tumour-mask gather ->
point-cloud transformer classifier.  Host code mirrors the reference's modules
(tfds_dense_descriptor, train_models, models_archs, create_pointcloud_dataframe,
visualization_utils, config_manager); all device arithmetic runs in libvdr.so through the
C ABI declared in include/vdr.h.  There is no CPU fallback."""

__version__ = "0.1.0"
