{
  "algorithm": "zlib",
  "chunk_bounds": [
    0,
    1000,
    2000
  ],
  "chunk_offsets": [
    0,
    62,
    129
  ],
  "chunk_order": "F",
  "comp_level": -1,
  "do_spatial_diff": false,
  "do_time_diff": true,
  "dtype": "int16",
  "n_channels": 16,
  "sample_rate": 1000.0,
  "sha1_compressed": "b90ae28d607754b5543e6454696bb2dcf93bf96e",
  "sha1_uncompressed": "d20122dc0f59ca0b42319652cf83bedbb05aabc0",
  "shape": [
    2000,
    16
  ],
  "version": "1.0"
}