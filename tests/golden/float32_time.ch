{
  "algorithm": "zlib",
  "chunk_bounds": [
    0,
    500,
    1000,
    1500
  ],
  "chunk_offsets": [
    0,
    13719,
    27459,
    41226
  ],
  "chunk_order": "F",
  "comp_level": -1,
  "do_spatial_diff": false,
  "do_time_diff": true,
  "dtype": "float32",
  "n_channels": 9,
  "sample_rate": 500.0,
  "sha1_compressed": "b1d1eb46a038d2b5d6b2e9a66e2db48f479f9587",
  "sha1_uncompressed": "1bc7a5d567a9b65f1661ce564e01448275c28a50",
  "shape": [
    1500,
    9
  ],
  "version": "1.0"
}