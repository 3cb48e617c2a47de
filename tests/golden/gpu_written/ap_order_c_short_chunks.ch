{
  "algorithm": "zlib",
  "chunk_bounds": [
    0,
    400,
    800,
    1200,
    1500
  ],
  "chunk_offsets": [
    0,
    7033,
    14063,
    21102,
    26466
  ],
  "chunk_order": "C",
  "comp_level": -1,
  "do_spatial_diff": false,
  "do_time_diff": true,
  "dtype": "int16",
  "n_channels": 20,
  "sample_rate": 1000.0,
  "sha1_compressed": "244db4627716038e2125cc0d53049a1a599b79e7",
  "sha1_uncompressed": "36b9d0a72f5f9e6c2d71ede3e82a200ee6bbfb44",
  "shape": [
    1500,
    20
  ],
  "version": "1.0"
}