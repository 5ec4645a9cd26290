-----------------------------------------------------------------------------
-- |
-- Module      :  Analysis.Parsimony.Dynamic.DirectOptimization.Pairwise.FFI.Batch
--
-- Binding of the batched entry points of @libpcgdo_b200.so@ (include/pcg_do_b200.h) for PCG.
-- Drop this file next to
-- lib/core/analysis/src/Analysis/Parsimony/Dynamic/DirectOptimization/Pairwise/FFI.hsc and add
-- @extra-libraries: pcgdo_b200@ to the @analysis@ library stanza (INTEGRATION.md §1).
--
-- NOT COMPILED in the build container of this repository (no GHC / hsc2hs there): the executable proof of the
-- ABI is the ctypes harness in tests/ (same struct layouts, same call sequence).  The marshalling below follows
-- the existing binding line by line (FFI.hsc:81-101 Storable Align_io, :222-260 algn2d, :474-515 buffers).
-----------------------------------------------------------------------------

{-# LANGUAGE ForeignFunctionInterface #-}
{-# LANGUAGE RecordWildCards          #-}

module Analysis.Parsimony.Dynamic.DirectOptimization.Pairwise.FFI.Batch
  ( PcgdoCtx
  , PcgdoTcm
  , withPcgdo
  , foreignPairwiseDOBatch
  ) where

import Control.Monad          (when)
import Data.Int
import Data.Word
import Foreign
import Foreign.C.String
import Foreign.C.Types

#include "pcg_do_b200.h"

data PcgdoCtx
data PcgdoTcm

-- | @pcgdo_batch@: one batch of raw align2d problems, one byte per element (alphabets <= 8).
data PcgdoBatch = PcgdoBatch
    { nPairs     :: CSize
    , elems      :: Ptr Word8
    , elemsBytes :: CSize
    , off1, off2 :: Ptr Word64
    , len1, len2 :: Ptr Word32
    , outOff     :: Ptr Word64
    , outBytes   :: CSize
    , outGapped, outSlot1, outSlot2 :: Ptr Word8
    , outLen     :: Ptr Word32
    , cost       :: Ptr Int32
    , cells      :: Ptr Word64
    , outOffActual :: Ptr Word64   -- ^ may be null; non-null = compact transfer: pair k's strings are at outOffActual[k]
    }

instance Storable PcgdoBatch where
    sizeOf    _ = (#size pcgdo_batch)
    alignment _ = alignment (undefined :: CSize)
    peek p = PcgdoBatch
        <$> (#peek pcgdo_batch, n_pairs)     p <*> (#peek pcgdo_batch, elems)      p
        <*> (#peek pcgdo_batch, elems_bytes) p
        <*> (#peek pcgdo_batch, off1)        p <*> (#peek pcgdo_batch, off2)       p
        <*> (#peek pcgdo_batch, len1)        p <*> (#peek pcgdo_batch, len2)       p
        <*> (#peek pcgdo_batch, out_off)     p <*> (#peek pcgdo_batch, out_bytes)  p
        <*> (#peek pcgdo_batch, out_gapped)  p <*> (#peek pcgdo_batch, out_slot1)  p
        <*> (#peek pcgdo_batch, out_slot2)   p <*> (#peek pcgdo_batch, out_len)    p
        <*> (#peek pcgdo_batch, cost)        p <*> (#peek pcgdo_batch, cells)      p
        <*> (#peek pcgdo_batch, out_off_actual) p
    poke p PcgdoBatch{..} = do
        (#poke pcgdo_batch, n_pairs)     p nPairs
        (#poke pcgdo_batch, elems)       p elems
        (#poke pcgdo_batch, elems_bytes) p elemsBytes
        (#poke pcgdo_batch, off1)        p off1
        (#poke pcgdo_batch, off2)        p off2
        (#poke pcgdo_batch, len1)        p len1
        (#poke pcgdo_batch, len2)        p len2
        (#poke pcgdo_batch, out_off)     p outOff
        (#poke pcgdo_batch, out_bytes)   p outBytes
        (#poke pcgdo_batch, out_gapped)  p outGapped
        (#poke pcgdo_batch, out_slot1)   p outSlot1
        (#poke pcgdo_batch, out_slot2)   p outSlot2
        (#poke pcgdo_batch, out_len)     p outLen
        (#poke pcgdo_batch, cost)        p cost
        (#poke pcgdo_batch, cells)       p cells
        (#poke pcgdo_batch, out_off_actual) p outOffActual

foreign import ccall unsafe "pcg_do_b200.h pcgdo_create"
    pcgdoCreate :: Ptr (Ptr PcgdoCtx) -> CInt -> IO CInt
foreign import ccall unsafe "pcg_do_b200.h pcgdo_destroy"
    pcgdoDestroy :: Ptr PcgdoCtx -> IO ()
foreign import ccall unsafe "pcg_do_b200.h pcgdo_last_error"
    pcgdoLastError :: Ptr PcgdoCtx -> IO CString
foreign import ccall unsafe "pcg_do_b200.h pcgdo_tcm_create"
    pcgdoTcmCreate :: Ptr PcgdoCtx -> Ptr CUInt -> CSize -> CUInt -> Ptr (Ptr PcgdoTcm) -> IO CInt
foreign import ccall unsafe "pcg_do_b200.h pcgdo_tcm_destroy"
    pcgdoTcmDestroy :: Ptr PcgdoCtx -> Ptr PcgdoTcm -> IO ()
-- 'safe': the call blocks for a whole batch (copies + kernels); it must not pin a capability.
foreign import ccall safe "pcg_do_b200.h pcgdo_align2d_batch_host"
    pcgdoAlign2dBatchHost :: Ptr PcgdoCtx -> Ptr PcgdoTcm -> Ptr PcgdoBatch -> IO CInt

-- | One context + one device TCM for the lifetime of the action.  @sigma@ is the a*a symbol-change matrix,
--   row-major, gap last — what 'generateDenseTransitionCostMatrix' (Data/TCM/Dense/FFI.hsc:347-377) passes to
--   @setUp2dCostMtx@.
withPcgdo :: Int -> [CUInt] -> Word -> ((Ptr PcgdoCtx, Ptr PcgdoTcm) -> IO a) -> IO a
withPcgdo device sigma a action =
    alloca $ \pCtx -> alloca $ \pTcm -> withArray sigma $ \pSigma -> do
        rc <- pcgdoCreate pCtx (toEnum device)
        when (rc /= 0) $ fail "pcgdo_create: no usable CUDA device (the library has no CPU fallback)"
        ctx <- peek pCtx
        rc' <- pcgdoTcmCreate ctx pSigma (toEnum (fromEnum a)) 0 pTcm
        when (rc' /= 0) $ pcgdoLastError ctx >>= peekCString >>= fail
        tcm <- peek pTcm
        r <- action (ctx, tcm)
        pcgdoTcmDestroy ctx tcm
        pcgdoDestroy ctx
        pure r

-- | @foreignPairwiseDO@ (FFI.hsc:179-188) over many pairs with ONE library call.
--   Input: for every pair the median strings of (longer, lesser) as produced by @measureAndUngapCharacters@
--   (Pairwise/Internal.hs:364-381), the LONGER first (FFI.hsc:261-268), one byte per element.
--   Output per pair: (cost, gapped medians, slot of input 1, slot of input 2) — @zip3@ them into
--   (median, left, right) triples exactly as @algn2d@ does at FFI.hsc:343, then @insertGaps@ / swap back.
foreignPairwiseDOBatch
  :: (Ptr PcgdoCtx, Ptr PcgdoTcm) -> [([Word8], [Word8])] -> IO [(Int, [Word8], [Word8], [Word8])]
foreignPairwiseDOBatch (ctx, tcm) pairs = do
    let n      = length pairs
        lens1  = fromIntegral . length . fst <$> pairs :: [Word32]
        lens2  = fromIntegral . length . snd <$> pairs :: [Word32]
        tot    = zipWith (\x y -> fromIntegral x + fromIntegral y) lens1 lens2 :: [Word64]
        offs1  = scanl (+) 0 tot
        offs2  = zipWith (\o l -> o + fromIntegral l) offs1 lens1
        caps   = (\t -> (t + 3) .&. complement 3) <$> tot
        oOffs  = scanl (+) 0 caps
        nBytes = fromIntegral (last offs1) + 16
        oBytes = fromIntegral (last oOffs) + 16
        flat   = concatMap (\(x, y) -> x <> y) pairs
    withArray (flat <> replicate 16 0) $ \pElems ->
      withArray (init offs1) $ \pOff1 -> withArray offs2 $ \pOff2 ->
      withArray lens1 $ \pLen1 -> withArray lens2 $ \pLen2 -> withArray (init oOffs) $ \pOOff ->
      allocaBytes oBytes $ \pG -> allocaBytes oBytes $ \pS1 -> allocaBytes oBytes $ \pS2 ->
      allocaArray n $ \pOLen -> allocaArray n $ \pCost -> allocaArray n $ \pActual ->
      with PcgdoBatch { nPairs = fromIntegral n, elems = pElems, elemsBytes = fromIntegral nBytes
                      , off1 = pOff1, off2 = pOff2, len1 = pLen1, len2 = pLen2
                      , outOff = pOOff, outBytes = fromIntegral oBytes
                      , outGapped = pG, outSlot1 = pS1, outSlot2 = pS2
                      , outLen = pOLen, cost = pCost, cells = nullPtr
                      , outOffActual = pActual } $ \pBatch -> do     -- compact transfer: only produced bytes come back
        rc <- pcgdoAlign2dBatchHost ctx tcm pBatch
        when (rc /= 0) $ pcgdoLastError ctx >>= peekCString >>= fail
        costs <- peekArray n pCost
        oLens <- peekArray n pOLen
        actual <- peekArray n pActual        -- where each pair's strings are (== init oOffs when nothing was packed)
        let slice p o l = peekArray (fromIntegral l) (p `plusPtr` fromIntegral o)
        sequence [ (,,,) (fromIntegral c) <$> slice pG o l <*> slice pS1 o l <*> slice pS2 o l
                 | (c, o, l) <- zip3 costs (actual :: [Word64]) oLens ]
