{
  "algorithm": "zlib",
  "chunk_bounds": [
    0,
    400,
    800,
    1200
  ],
  "chunk_offsets": [
    0,
    10789,
    21573,
    32360
  ],
  "chunk_order": "F",
  "comp_level": -1,
  "do_spatial_diff": true,
  "do_time_diff": true,
  "dtype": "float32",
  "n_channels": 7,
  "sample_rate": 400.0,
  "sha1_compressed": "5918afa0a3d6c2a3aa0b2aa717396c5059947dda",
  "sha1_uncompressed": "cb2e0020f51d4febe2b161f56b9eeac05d72c286",
  "shape": [
    1200,
    7
  ],
  "version": "1.0"
}