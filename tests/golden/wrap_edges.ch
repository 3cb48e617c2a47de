{
  "algorithm": "zlib",
  "chunk_bounds": [
    0,
    30,
    60,
    90,
    100
  ],
  "chunk_offsets": [
    0,
    49,
    95,
    144,
    184
  ],
  "chunk_order": "F",
  "comp_level": -1,
  "do_spatial_diff": false,
  "do_time_diff": true,
  "dtype": "int16",
  "n_channels": 15,
  "sample_rate": 30.0,
  "sha1_compressed": "42f77bf56fdce9f6a7b794572be54186f69c974b",
  "sha1_uncompressed": "eadf179b686dda54a75cd8677bc94ea76c380387",
  "shape": [
    100,
    15
  ],
  "version": "1.0"
}