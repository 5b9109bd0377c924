{-# LANGUAGE ForeignFunctionInterface #-}
-- | ECC.Code.LDPC.GPU.B200 -- binding of libldpc_b200.so into ku-fpg/ecc-ldpc.
--
-- Source only: this image has no GHC, so the module is exercised by review, and the same contract is
-- proven through the ctypes mirror (ecc-ldpc_b200/ecc.py) in tests/test_gpu_parity.py::test_ecc_record_mirror.
--
-- It registers five decoders next to the existing ECC.Code.LDPC.* ones, built with the reference's own
-- 'mkLDPC_CodeIO' (src/ECC/Code/LDPC/Utils.hs:91-108) exactly like GPU/CUDA/Arraylet2.hs:60-61 does:
--
--   ldpc/b200-minsum/<matrix>/<max-rounds>[/<num>/<den>]      fp32, normalised min-sum (alpha = 3/4)
--   ldpc/b200-spa/<matrix>/<max-rounds>[/<num>/<den>]         fp32, sum-product
--   ldpc/b200-minsum-fp64 , ldpc/b200-spa-fp64                 fp64 verification instances
--   ldpc/b200-spa-cm-fp64                                      the cached-product variant of Fast/CachedMult.hs (fp64)
--
-- main/Main.hs only needs  `import qualified ECC.Code.LDPC.GPU.B200 as B200`  and  `<> B200.code`  in `codes`.
module ECC.Code.LDPC.GPU.B200 (code) where

import ECC.Code.LDPC.Utils
import ECC.Types
import qualified ECC.Code.LDPC.Fast.Encoder as E
import qualified Data.Matrix.QuasiCyclic as Q
import Data.Matrix (toLists)
import Data.Bits (popCount, testBit)
import Data.Monoid ((<>))
import qualified Data.Vector.Unboxed as U
import qualified Data.Vector.Storable as S
import Foreign
import Foreign.C.Types
import Foreign.C.String

data CodeH
data DecH

foreign import ccall safe "ldpc_b200_init"            c_init    :: CInt -> IO CInt
foreign import ccall safe "ldpc_b200_shutdown"        c_fini    :: IO ()
foreign import ccall safe "ldpc_b200_last_error"      c_err     :: IO CString
foreign import ccall safe "ldpc_b200_code_create_qc"  c_code_qc :: CInt -> CInt -> CInt -> Ptr Int32 -> Ptr (Ptr CodeH) -> IO CInt
foreign import ccall safe "ldpc_b200_decoder_create"  c_dec_new :: Ptr CodeH -> CInt -> CInt -> CInt -> Ptr (Ptr DecH) -> IO CInt
foreign import ccall safe "ldpc_b200_decode"          c_decode  :: Ptr DecH -> CInt -> CInt -> Ptr Double -> Ptr Word8 -> Ptr Int32 -> Ptr Word8 -> IO CInt

check :: String -> CInt -> IO ()
check what rc | rc == 0   = return ()
              | otherwise = do msg <- peekCString =<< c_err
                               error (what ++ ": libldpc_b200 error " ++ show rc ++ ": " ++ msg)

-- | Circulant offsets in the form of GPU/CUDA/Arraylet2.hs:325-331: bit index of the single set bit, -1 for 0.
offsets :: Q.QuasiCyclic Integer -> [Int32]
offsets (Q.QuasiCyclic _ m) = map f (concat (toLists m))
  where f 0 = -1
        f n | popCount n == 1 = fromIntegral (length (takeWhile (not . testBit n) [0..]))
            | otherwise       = error ("QuasiCyclic matrix has non-powers of two initial value of " ++ show n)

-- | One `decoder0 h` call (Utils.hs:53): device copy of H + a decoder handle with its own stream.
decoder :: CInt -> CInt -> () -> Q.QuasiCyclic Integer -> IO (Rate -> Int -> U.Vector Double -> IO (Maybe (U.Vector Bool)))
decoder algo prec () h@(Q.QuasiCyclic sz m) = do
  let offs = S.fromList (offsets h)
      (br, bc) = (length (toLists m), length (head (toLists m)))
      n = sz * bc
  codeH <- alloca $ \out -> do
             S.unsafeWith offs $ \p -> check "code_create_qc" =<< c_code_qc (fromIntegral sz) (fromIntegral br) (fromIntegral bc) p out
             peek out
  decH  <- alloca $ \out -> do
             check "decoder_create" =<< c_dec_new codeH algo prec 1 out
             peek out
  return $ \ _rate maxIterations orig_lam -> do
    let lam = S.convert orig_lam :: S.Vector Double          -- Arraylet2.hs:63-67 pattern
    bits <- mallocForeignPtrBytes n
    S.unsafeWith lam $ \pl -> withForeignPtr bits $ \pb ->
      check "decode" =<< c_decode decH (fromIntegral maxIterations) 1 pl pb nullPtr nullPtr
    let out = S.unsafeFromForeignPtr0 bits n :: S.Vector Word8
    return $ Just $ U.convert $ S.map (/= 0) out

-- | maxThreadCount = 16: mkLDPC calls `decoder0 h` that many times up front (Utils.hs:53) and hands tester thread t the
-- decoder (t mod 16) (Utils.hs:69).  A decoder handle owns its stream and staging buffers and is not re-entrant, so one
-- handle per tester thread is what makes concurrent `decode` calls safe without a lock (the reference's own CUDA
-- variants pass 1 and share their buffers, GPU/CUDA/Arraylet2.hs:61).
mk :: String -> CInt -> CInt -> Code
mk name algo prec = mkLDPC_CodeIO name 16 E.encoder (decoder algo prec)
                                  (check "init" =<< c_init (-1)) (const c_fini)

code :: Code
code = mk "b200-minsum" 1 1 <> mk "b200-spa" 0 1 <> mk "b200-minsum-fp64" 1 0 <> mk "b200-spa-fp64" 0 0
    <> mk "b200-spa-cm-fp64" 2 0
