{
  "algorithm": "zlib",
  "chunk_bounds": [
    0,
    1000,
    2000,
    2500
  ],
  "chunk_offsets": [
    0,
    38196,
    76292,
    95887
  ],
  "chunk_order": "F",
  "comp_level": -1,
  "do_spatial_diff": true,
  "do_time_diff": true,
  "dtype": "int16",
  "n_channels": 40,
  "sample_rate": 1000.0,
  "sha1_compressed": "11bdccd57327fda9046c2f70d733ad185dd1a76c",
  "sha1_uncompressed": "3e219b584766b271cc7b4da961d9d4f7a296cddb",
  "shape": [
    2500,
    40
  ],
  "version": "1.0"
}