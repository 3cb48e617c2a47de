{
  "algorithm": "zlib",
  "chunk_bounds": [
    0,
    300,
    600,
    700
  ],
  "chunk_offsets": [
    0,
    87830,
    176159,
    206239
  ],
  "chunk_order": "F",
  "comp_level": -1,
  "do_spatial_diff": false,
  "do_time_diff": true,
  "dtype": "int16",
  "n_channels": 385,
  "sample_rate": 300.0,
  "sha1_compressed": "aaea3eceb32d5beefdaf3651654961f693b51e56",
  "sha1_uncompressed": "bc32230c0a44299431595b409d0fc48e0d77f1a3",
  "shape": [
    700,
    385
  ],
  "version": "1.0"
}