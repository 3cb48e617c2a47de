{
  "algorithm": "zlib",
  "chunk_bounds": [
    0,
    100,
    200,
    300
  ],
  "chunk_offsets": [
    0,
    2166,
    4335,
    6482
  ],
  "chunk_order": "F",
  "comp_level": -1,
  "do_spatial_diff": false,
  "do_time_diff": false,
  "dtype": "int16",
  "n_channels": 24,
  "sample_rate": 100.0,
  "sha1_compressed": "55488a005fa3bd367a15c6dd944b1ff2702dc67c",
  "sha1_uncompressed": "ecc103d6b4b689b01bff8d3fff8909f07618131b",
  "shape": [
    300,
    24
  ],
  "version": "1.0"
}