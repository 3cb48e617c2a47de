{
  "algorithm": "zlib",
  "chunk_bounds": [
    0,
    200,
    400,
    500
  ],
  "chunk_offsets": [
    0,
    11082,
    22133,
    27852
  ],
  "chunk_order": "C",
  "comp_level": -1,
  "do_spatial_diff": false,
  "do_time_diff": true,
  "dtype": "int16",
  "n_channels": 64,
  "sample_rate": 200.0,
  "sha1_compressed": "1e8d02d8029230fada39d8e8230699bd336eb725",
  "sha1_uncompressed": "58fdce8d1e9d2c670c10499e7c2c5e826721aeca",
  "shape": [
    500,
    64
  ],
  "version": "1.0"
}