//! dump_golden -- run the REAL akaze crate on an image and write everything the B200 repo needs to pin its
//! oracle against the reference (SURVEY.md section 7 "Hard parts").
//!
//! No Rust toolchain exists in the build image, so this file is shipped uncompiled. On any machine with cargo:
//!
//!     git clone https://github.com/indianajohn/akaze-rust && cd akaze-rust
//!     cp /path/to/repo/tools/dump_golden.rs akaze-util/src/bin/dump_golden.rs
//!     cargo run --release --bin dump_golden -- test-data/1.jpg reference_1
//!     cargo run --release --bin dump_golden -- test-data/2.jpg reference_2 reference_1.bin
//!     cp reference_* /path/to/repo/tests/golden/
//!
//! and `python -m pytest tests/test_reference_goldens.py` stops skipping. Outputs for prefix P:
//!   P.luma   u32 width, u32 height (LE), then width*height bytes: the gray image exactly as the `image` crate decoded
//!            it and `to_luma` converted it (the engine's input boundary; removes JPEG-decoder differences)
//!   P.bin    akaze_util::serialize_features_to_file (bincode 1.1): keypoints + descriptors
//!   P.evo    u32 n_levels, then per level: u32 width, u32 height, Lt (w*h f32 LE), Ldet (w*h f32 LE)
//!   P.meta   one line of text: contrast-independent run facts (keypoint count, level count)
//! With a third argument (another image's P.bin): P.matches.bin = descriptor_match(that image, this image, 10000,
//! 0.86) as akaze_util::serialize_matches_to_file writes it (the deterministic first half of akaze::match_features,
//! akaze/src/lib.rs:261-266; RANSAC is seeded from the clock and not reproducible).
use akaze::ops::feature_matching::descriptor_match;
use akaze::types::evolution::Config;
use akaze::types::image::ImageFunctions;
use akaze_util::{deserialize_features_from_file, serialize_features_to_file, serialize_matches_to_file, Features};
use std::fs::File;
use std::io::Write;
use std::path::PathBuf;

fn main() {
    let args: Vec<String> = std::env::args().collect();
    if args.len() < 3 {
        eprintln!("usage: dump_golden IMAGE PREFIX [OTHER_FEATURES.bin]");
        std::process::exit(2);
    }
    let (image_path, prefix) = (PathBuf::from(&args[1]), args[2].clone());

    // the gray image the crate sees (akaze/src/lib.rs:171, types/image.rs:128)
    let gray = image::open(&image_path).unwrap().to_luma();
    let mut f = File::create(format!("{}.luma", prefix)).unwrap();
    f.write_all(&(gray.width() as u32).to_le_bytes()).unwrap();
    f.write_all(&(gray.height() as u32).to_le_bytes()).unwrap();
    f.write_all(&gray.clone().into_raw()).unwrap();

    let (evolutions, keypoints, descriptors) = akaze::extract_features(image_path, Config::default());

    let mut f = File::create(format!("{}.evo", prefix)).unwrap();
    f.write_all(&(evolutions.len() as u32).to_le_bytes()).unwrap();
    for e in &evolutions {
        f.write_all(&(e.Lt.width() as u32).to_le_bytes()).unwrap();
        f.write_all(&(e.Lt.height() as u32).to_le_bytes()).unwrap();
        for img in &[&e.Lt, &e.Ldet] {
            for v in &img.buffer {
                f.write_all(&v.to_le_bytes()).unwrap();
            }
        }
    }
    let mut f = File::create(format!("{}.meta", prefix)).unwrap();
    writeln!(f, "keypoints={} levels={}", keypoints.len(), evolutions.len()).unwrap();

    let features = Features { keypoints, descriptors };
    serialize_features_to_file(&features, format!("{}.bin", prefix)).unwrap();

    if args.len() > 3 {
        let other = deserialize_features_from_file(&args[3]).unwrap();
        let m = descriptor_match(&other.descriptors, &features.descriptors, 10000, 0.86);
        serialize_matches_to_file(&m, format!("{}.matches.bin", prefix)).unwrap();
    }
}
