{-# LANGUAGE ForeignFunctionInterface #-}
-- | ECC.Code.LDPC.GPU.B200 -- binding of libldpc_b200.so into ku-fpg/ecc-ldpc.
--
-- Source only: this image has no GHC, so the module is exercised by review, and the same contract is proven through
-- the C++ host layer (ldpc_b200_ecc_*) and its ctypes mirror in tests/test_gpu_parity.py (test_ecc_record_mirror,
-- test_tester_pattern_harness).
--
-- It registers the decoders next to the existing ECC.Code.LDPC.* ones, built with the reference's own 'mkLDPC_CodeIO'
-- (src/ECC/Code/LDPC/Utils.hs:91-108) exactly like GPU/CUDA/Arraylet2.hs:60-61 does:
--
--   ldpc/b200-minsum/<matrix>/<max-rounds>[/<num>/<den>]      fp32, normalised min-sum (alpha = 3/4)
--   ldpc/b200-spa/<matrix>/<max-rounds>[/<num>/<den>]         fp32, sum-product
--   ldpc/b200-minsum-fp64 , ldpc/b200-spa-fp64                 fp64 verification instances
--   ldpc/b200-spa-cm-fp64                                      the cached-product variant of Fast/CachedMult.hs (fp64)
--   ldpc/b200-arraylet-fp64 , ldpc/b200-arraylet-min-fp64      the arraylet association of Fast/Arraylet{,Min}.hs (fp64)
--
-- ANY matrix the reference can load works under these names: `.q` files arrive as 'QuasiCyclic Integer' and keep their
-- circulant form (ldpc_b200_code_create_qc; the compiled RESIDENT kernels for the JPL codes), `.alist` / `.m` files
-- arrive as 'Matrix Bool' (Data/BitMatrix/Loader.hs:34-39) and go through ldpc_b200_code_create_csr -- so
-- `ldpc/b200-spa/moon.7.13/20`, BASELINE config 1, loads (precedent for a 'Matrix Bool' decoder: Reference/Orig.hs:20-31).
--
-- main/Main.hs only needs  `import qualified ECC.Code.LDPC.GPU.B200 as B200`  and  `<> B200.code`  in `codes`.
module ECC.Code.LDPC.GPU.B200 (code) where

import ECC.Code.LDPC.Utils
import ECC.Types
import qualified ECC.Code.LDPC.Fast.Encoder as E
import qualified Data.Matrix.QuasiCyclic as Q
import Data.BitMatrix.Loader
import Data.Matrix (Matrix, toLists, nrows, ncols, getCol)
import Data.Bits (popCount, testBit, xor)
import Data.IORef
import Data.List (foldl')
import Data.Monoid ((<>))
import qualified Data.Vector as V
import qualified Data.Vector.Unboxed as U
import qualified Data.Vector.Storable as S
import Foreign
import Foreign.C.Types
import Foreign.C.String

data CodeH
data DecH

foreign import ccall safe "ldpc_b200_init"             c_init     :: CInt -> IO CInt
foreign import ccall safe "ldpc_b200_shutdown"         c_fini     :: IO ()
foreign import ccall safe "ldpc_b200_last_error"       c_err      :: IO CString
foreign import ccall safe "ldpc_b200_code_create_qc"   c_code_qc  :: CInt -> CInt -> CInt -> Ptr Int32 -> Ptr (Ptr CodeH) -> IO CInt
foreign import ccall safe "ldpc_b200_code_create_csr"  c_code_csr :: CInt -> CInt -> Ptr Int32 -> Ptr Int32 -> Ptr (Ptr CodeH) -> IO CInt
foreign import ccall safe "ldpc_b200_code_destroy"     c_code_del :: Ptr CodeH -> IO ()
foreign import ccall safe "ldpc_b200_decoder_create"   c_dec_new  :: Ptr CodeH -> CInt -> CInt -> CInt -> Ptr (Ptr DecH) -> IO CInt
foreign import ccall safe "ldpc_b200_decoder_destroy"  c_dec_del  :: Ptr DecH -> IO ()
foreign import ccall safe "ldpc_b200_decode"           c_decode   :: Ptr DecH -> CInt -> CInt -> Ptr Double -> Ptr Word8 -> Ptr Int32 -> Ptr Word8 -> IO CInt

check :: String -> CInt -> IO ()
check what rc | rc == 0   = return ()
              | otherwise = do msg <- peekCString =<< c_err
                               error (what ++ ": libldpc_b200 error " ++ show rc ++ ": " ++ msg)

-- | A matrix in whatever form its file has: the circulant description of a `.q` file, the dense bit matrix of an
-- `.alist` / `.m` file.  'loadMatrix' (Loader.hs:64-81) takes the first file whose type converts; this instance
-- converts all three (the reference's own two instances are Loader.hs:34-46).
data AnyMatrix = AnyQC (Q.QuasiCyclic Integer) | AnyDense (Matrix Bool)

instance MatrixLoader AnyMatrix where
  getMatrix (LoaderQC m)     = pure (AnyQC m)
  getMatrix (LoaderAline m)  = pure (AnyDense m)
  getMatrix (LoaderMatlab m) = pure (AnyDense m)
  getNRows (AnyQC m)    = getNRows m
  getNRows (AnyDense m) = nrows m
  getNCols (AnyQC m)    = getNCols m
  getNCols (AnyDense m) = ncols m

-- | Circulant offsets in the form of GPU/CUDA/Arraylet2.hs:325-331: bit index of the single set bit, -1 for 0.
-- 'Nothing' when some block carries more than one circulant (e.g. the dense generator blocks): such a matrix is handed
-- over expanded.
offsets :: Q.QuasiCyclic Integer -> Maybe [Int32]
offsets (Q.QuasiCyclic _ m) = mapM f (concat (toLists m))
  where f 0 = Just (-1)
        f n | popCount n == 1 = Just (fromIntegral (length (takeWhile (not . testBit n) [0..])))
            | otherwise       = Nothing

-- | CSR of a dense bit matrix: row pointers and 0-based column indices, ascending inside a row -- the order in which
-- Reference/Orig.hs:82-92 visits the ones of H.
csr :: Matrix Bool -> ([Int32], [Int32])
csr m = (scanl (+) 0 (map (fromIntegral . length) cols), concat cols)
  where cols = [ [ j | (j, True) <- zip [0 ..] row ] | row <- toLists m ]

-- | Device copy of H.  Identical matrices share one device object inside the library, so the `maxThreadCount` calls of
-- 'decoder' below (Utils.hs:53) cost one copy and all their decoder handles meet in one mailbox.
createCode :: AnyMatrix -> IO (Ptr CodeH)
createCode (AnyQC h@(Q.QuasiCyclic sz m))
  | Just offs <- offsets h = alloca $ \out -> do
      S.unsafeWith (S.fromList offs) $ \p ->
        check "code_create_qc" =<< c_code_qc (fromIntegral sz) (fromIntegral (nrows m)) (fromIntegral (ncols m)) p out
      peek out
  | otherwise = createCode (AnyDense (Q.toBitMatrix h))
createCode (AnyDense m) = alloca $ \out -> do
  let (rp, ci) = csr m
  S.unsafeWith (S.fromList rp) $ \prp -> S.unsafeWith (S.fromList (if null ci then [0] else ci)) $ \pci ->
    check "code_create_csr" =<< c_code_csr (fromIntegral (nrows m)) (fromIntegral (ncols m)) prp pci out
  peek out

-- | What 'initialize' hands to every 'decoder' call and to 'finalize': the destructors of everything created so far.
type Vars = IORef [IO ()]

-- | One `decoder0 h` call (Utils.hs:53): (shared) device copy of H + a decoder handle of its own.
decoder :: CInt -> CInt -> Vars -> AnyMatrix -> IO (Rate -> Int -> U.Vector Double -> IO (Maybe (U.Vector Bool)))
decoder algo prec vars h = do
  let n = getNCols h
  codeH <- createCode h
  decH  <- alloca $ \out -> do
             check "decoder_create" =<< c_dec_new codeH algo prec 1 out
             peek out
  modifyIORef' vars ([c_dec_del decH, c_code_del codeH] ++)       -- decoders before their codes
  return $ \ _rate maxIterations orig_lam -> do
    let lam = S.convert orig_lam :: S.Vector Double          -- Arraylet2.hs:63-67 pattern
    bits <- mallocForeignPtrBytes n
    -- one frame per call: inside the library this goes through the mailbox shared by the handles of this code (no copies,
    -- no stream synchronisation; frames of concurrently calling tester threads are coalesced into one launch)
    S.unsafeWith lam $ \pl -> withForeignPtr bits $ \pb ->
      check "decode" =<< c_decode decH (fromIntegral maxIterations) 1 pl pb nullPtr nullPtr
    let out = S.unsafeFromForeignPtr0 bits n :: S.Vector Word8
    return $ Just $ U.convert $ S.map (/= 0) out

-- | `encode` (Utils.hs:61 takes the parity part): the circulant word multiply of Fast/Encoder.hs:26-63 for `.q`
-- generators, v * G over GF(2) (Reference/Orig.hs:25-26) for dense ones.
encoder :: AnyMatrix -> Rate -> U.Vector Bool -> U.Vector Bool
encoder (AnyQC g)    r v = E.encoder g r v
encoder (AnyDense g) _ v = U.fromList [ foldl' xor False [ gij | (True, gij) <- zip vs (V.toList (getCol j g)) ] | j <- [1 .. ncols g] ]
  where vs = U.toList v

-- | maxThreadCount = 16: mkLDPC calls `decoder0 h` that many times up front (Utils.hs:53) and hands tester thread t the
-- decoder (t mod 16) (Utils.hs:69).  A decoder handle is not re-entrant, so one handle per tester thread is what makes
-- concurrent `decode` calls safe without a lock (the reference's own CUDA variants pass 1 and share their buffers,
-- GPU/CUDA/Arraylet2.hs:61).  `finalize` destroys the handles and then shuts the library down (which would also release
-- anything still alive).
mk :: String -> CInt -> CInt -> Code
mk name algo prec = mkLDPC_CodeIO name 16 encoder (decoder algo prec)
                                  (do check "init" =<< c_init (-1); newIORef [])
                                  (\vars -> do sequence_ =<< readIORef vars; writeIORef vars []; c_fini)

code :: Code
code = mk "b200-minsum" 1 1 <> mk "b200-spa" 0 1 <> mk "b200-minsum-fp64" 1 0 <> mk "b200-spa-fp64" 0 0
    <> mk "b200-spa-cm-fp64" 2 0 <> mk "b200-arraylet-fp64" 3 0 <> mk "b200-arraylet-min-fp64" 4 0
