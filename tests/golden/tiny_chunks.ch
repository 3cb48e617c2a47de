{
  "algorithm": "zlib",
  "chunk_bounds": [
    0,
    12,
    24,
    36,
    48,
    60,
    72,
    84,
    96,
    100
  ],
  "chunk_offsets": [
    0,
    407,
    812,
    1223,
    1632,
    2037,
    2444,
    2848,
    3253,
    3409
  ],
  "chunk_order": "F",
  "comp_level": -1,
  "do_spatial_diff": false,
  "do_time_diff": true,
  "dtype": "int16",
  "n_channels": 19,
  "sample_rate": 1234.0,
  "sha1_compressed": "764489a8f03a497cc5742cf12ed65ff4176fffbe",
  "sha1_uncompressed": "e9b53e08f47caea8ba890b93c193f29b9e0e0cbe",
  "shape": [
    100,
    19
  ],
  "version": "1.0"
}