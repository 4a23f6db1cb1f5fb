//! The reference's example binary (`src/main.rs:9-34`) against this crate: message `[1; 16]`, key `[0; 16]`.
use aes::cipher::{generic_array::GenericArray, BlockEncrypt, KeyInit};
use aes::Aes128;
use zk_aes::{encrypt, synthesize_keys, verify_encryption};

fn main() -> anyhow::Result<()> {
    let message = [1_u8; 16];
    let secret_key = [0_u8; 16];
    let mut block = GenericArray::clone_from_slice(&message);
    Aes128::new(GenericArray::from_slice(&secret_key)).encrypt_block(&mut block);
    let primitive_ciphertext = block.to_vec();

    let (proving_key, verifying_key) = synthesize_keys(message.len())?;
    let proof = encrypt(&message, &secret_key, proving_key)?;
    assert_eq!(proof.ciphertext(), primitive_ciphertext.as_slice());
    let result = verify_encryption(verifying_key, &proof, &primitive_ciphertext)?;
    assert!(result);
    println!("proof of {} bytes verified", proof.as_bytes().len());
    Ok(())
}
