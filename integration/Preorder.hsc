-----------------------------------------------------------------------------
-- |
-- Module      :  Analysis.Parsimony.Dynamic.DirectOptimization.Pairwise.FFI.Preorder
--
-- Binding of the pre-order entry points of @libpcgdo_b200.so@ (include/pcg_do_b200.h, "pre-order pass") for PCG:
-- the final (implied) alignments of every node of a rooted binary tree and one dynamic character from the post-order
-- alignment contexts — what @directOptimizationPreorder@ (DirectOptimization/Internal.hs:148-286) computes node by
-- node under @preorderFromRooting@ (Bio/Graph/PhylogeneticDAG/Preorder.hs:359-520) — in one call.
--
-- NOT COMPILED in the build container of this repository (no GHC / hsc2hs there): the executable model of this
-- call sequence is pcg_b200/preorder.py (@preorder_device@, @preorder_from_rooting@), driven by
-- tests/test_gpu_forest.py through the same C entry points.
-----------------------------------------------------------------------------

{-# LANGUAGE ForeignFunctionInterface #-}

module Analysis.Parsimony.Dynamic.DirectOptimization.Pairwise.FFI.Preorder
  ( impliedAlignments
  ) where

import Control.Monad    (forM, when)
import Data.Int
import Data.Word
import Foreign
import Foreign.C.String
import Foreign.C.Types

#include "pcg_do_b200.h"

data PcgdoCtx
data PcgdoPreorder

foreign import ccall unsafe "pcg_do_b200.h pcgdo_last_error"
    pcgdoLastError :: Ptr PcgdoCtx -> IO CString
foreign import ccall unsafe "pcg_do_b200.h pcgdo_preorder_create"
    pcgdoPreorderCreate :: Ptr PcgdoCtx -> Word32 -> Word32 -> Word32 -> Ptr Int32 -> Ptr (Ptr PcgdoPreorder) -> IO CInt
foreign import ccall unsafe "pcg_do_b200.h pcgdo_preorder_destroy"
    pcgdoPreorderDestroy :: Ptr PcgdoCtx -> Ptr PcgdoPreorder -> IO ()
foreign import ccall unsafe "pcg_do_b200.h pcgdo_preorder_set_contexts"
    pcgdoPreorderSetContexts :: Ptr PcgdoCtx -> Ptr PcgdoPreorder -> Ptr Word8 -> Ptr Word8 -> Ptr Word8 -> CSize
                             -> Ptr Word64 -> Ptr Word32 -> IO CInt
-- `safe`: one launch per tree depth plus a synchronisation; do not hold a capability meanwhile
foreign import ccall safe   "pcg_do_b200.h pcgdo_preorder_run"
    pcgdoPreorderRun :: Ptr PcgdoCtx -> Ptr PcgdoPreorder -> CInt -> Ptr () -> IO CInt
foreign import ccall unsafe "pcg_do_b200.h pcgdo_preorder_length"
    pcgdoPreorderLength :: Ptr PcgdoPreorder -> Word32 -> Word32 -> IO Word32
foreign import ccall unsafe "pcg_do_b200.h pcgdo_preorder_get"
    pcgdoPreorderGet :: Ptr PcgdoCtx -> Ptr PcgdoPreorder -> Word32 -> Word32 -> Word32
                     -> Ptr Word8 -> Ptr Word8 -> Ptr Word8 -> Ptr Word8 -> Ptr CInt -> IO CInt

missingLength :: Word32
missingLength = (#const PCGDO_MISSING)

-- | One tree, one character.
--   @children@: (left, right) node ids of the internal nodes @nTaxa ..@, leaves are @0 .. nTaxa - 1@ — for the
--   re-rooted traversal focus of a character pass the RE-ROOTED tree (the virtual root takes the original root's id,
--   every other node lists the two neighbours it was not entered from, INTEGRATION.md §2d).
--   @contexts@: per node @Nothing@ (Missing) or the @(median, left, right)@ triples of its post-order decoration
--   (@^. alignmentContext@; a leaf: @(s, s, s)@), one byte per ambiguity set.
--   Result per node: @Nothing@, or (final alignment triples, single disambiguation) — @impliedAlignment@ and
--   @singleDisambiguation@ of 'extendPostorderToDirectOptimization'.
impliedAlignments
  :: Ptr PcgdoCtx -> Int -> Int -> [(Int32, Int32)] -> [Maybe [(Word8, Word8, Word8)]]
  -> IO [Maybe ([(Word8, Word8, Word8)], [Word8])]
impliedAlignments ctx alphabetSize nTaxa children contexts = do
    let nNodes = 2 * nTaxa - 1
        lens   = maybe missingLength (fromIntegral . length) <$> contexts :: [Word32]
        sizes  = maybe 0 (fromIntegral . length) <$> contexts            :: [Word64]
        offs   = init (scanl (+) 0 sizes)
        flat   = concatMap (maybe [] id) contexts
        total  = length flat
        m      = (\(x, _, _) -> x) <$> flat
        l      = (\(_, x, _) -> x) <$> flat
        r      = (\(_, _, x) -> x) <$> flat
        failWith = pcgdoLastError ctx >>= peekCString >>= fail
    alloca $ \pPre -> withArray (concatMap (\(a, b) -> [a, b]) children) $ \pChildren -> do
        rc <- pcgdoPreorderCreate ctx 1 (fromIntegral nTaxa) 1 pChildren pPre
        when (rc /= 0) failWith
        pre <- peek pPre
        withArray (m <> [0]) $ \pM -> withArray (l <> [0]) $ \pL -> withArray (r <> [0]) $ \pR ->
          withArray offs $ \pOff -> withArray lens $ \pLen -> do
            rc1 <- pcgdoPreorderSetContexts ctx pre pM pL pR (fromIntegral total) pOff pLen
            when (rc1 /= 0) failWith
        -- PCGDO_ERR_OVERFLOW here is the reference's "Impossible happened in 'deriveImpliedAlignment'"
        rc2 <- pcgdoPreorderRun ctx pre (fromIntegral alphabetSize) nullPtr
        when (rc2 /= 0) failWith
        n <- fromIntegral <$> pcgdoPreorderLength pre 0 0
        out <- forM [0 .. nNodes - 1] $ \v ->
            if n == fromIntegral missingLength then pure Nothing else
            allocaArray n $ \qM -> allocaArray n $ \qL -> allocaArray n $ \qR -> allocaArray n $ \qS ->
            alloca $ \pMissing -> do
                rc3 <- pcgdoPreorderGet ctx pre 0 0 (fromIntegral v) qM qL qR qS pMissing
                when (rc3 /= 0) failWith
                missing <- peek pMissing
                if missing /= 0 then pure Nothing else do
                    ms <- peekArray n qM; ls <- peekArray n qL; rs <- peekArray n qR; ss <- peekArray n qS
                    pure $ Just (zip3 ms ls rs, ss)
        pcgdoPreorderDestroy ctx pre
        pure out
