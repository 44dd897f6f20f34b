package io;

import org.apache.log4j.Logger;
import ru.ifmo.genetics.dna.Dna;
import ru.ifmo.genetics.dna.DnaTools;
import ru.ifmo.genetics.io.ReadersUtils;
import ru.ifmo.genetics.io.sources.NamedSource;
import ru.ifmo.genetics.statistics.QuickQuantitativeStatistics;
import ru.ifmo.genetics.utils.tool.ExecutionFailedException;

import java.io.*;
import java.lang.foreign.*;

import static java.lang.foreign.ValueLayout.*;

/**
 * Drop-in for the two calls KmersCounterMain.runImpl makes (src/tools/KmersCounterMain.java:77-78,99):
 *   IOUtils.loadReads(files, k, 0, P, logger)  +  IOUtils.printKmers(hm, b, outFile, stFile)
 * The ITMO readers keep doing the parsing (so BINQ / bz2 / IUPAC behave exactly as before); the batches of
 * ReadsDispatcher (<= 32768 reads, src/io/ReadsDispatcher.java:34-53) are written as ASCII into a pinned
 * buffer and handed to the GPU.  NOT compiled here (no JDK); see INTEGRATION.md.
 */
public final class GpuKmerCounting {
    public static long countAndPrint(File[] files, int k, int threshold, File outFile, File stFile, Logger logger)
            throws ExecutionFailedException, IOException {
        try (Arena arena = Arena.ofConfined()) {
            MemorySegment cfg = arena.allocate(MfkcNative.CFG);
            cfg.set(JAVA_INT, 0, (int) MfkcNative.CFG.byteSize());
            cfg.set(JAVA_INT, 4, k);
            long bases = 0;
            for (File f : files) bases += f.length();
            cfg.set(JAVA_LONG, MfkcNative.CFG.byteOffset(MemoryLayout.PathElement.groupElement("expected_kmers")), bases);
            MemorySegment pctx = arena.allocate(ADDRESS);
            int rc = (int) MfkcNative.CREATE.invokeExact(cfg, pctx);
            MfkcNative.check(MemorySegment.NULL, rc);
            MemorySegment ctx = pctx.get(ADDRESS, 0);
            try {
                final int capReads = 1 << 20;
                final long capBases = 256L << 20;
                MemorySegment pp = arena.allocate(ADDRESS);
                MfkcNative.check(ctx, (int) MfkcNative.PINNED_ALLOC.invokeExact(ctx, capBases, pp));
                MemorySegment hBases = pp.get(ADDRESS, 0).reinterpret(capBases);
                MfkcNative.check(ctx, (int) MfkcNative.PINNED_ALLOC.invokeExact(ctx, 8L * (capReads + 1), pp));
                MemorySegment hOffs = pp.get(ADDRESS, 0).reinterpret(8L * (capReads + 1));
                for (File file : files) {                               // IOUtils.run(files, workers, ..) :838-865
                    NamedSource<Dna> reader = ReadersUtils.readDnaLazy(file);
                    int n = 0; long used = 0;
                    hOffs.set(JAVA_LONG, 0, 0L);
                    for (Dna dna : reader) {
                        if (n == capReads || used + dna.length() > capBases) {
                            MfkcNative.check(ctx, (int) MfkcNative.SUBMIT_READS.invokeExact(ctx, hBases, hOffs, n));
                            n = 0; used = 0;
                        }
                        for (int i = 0; i < dna.length(); i++)
                            hBases.set(JAVA_BYTE, used + i, (byte) DnaTools.toChar(dna.nucAt(i)));
                        used += dna.length();
                        hOffs.set(JAVA_LONG, 8L * (++n), used);
                    }
                    if (n > 0) MfkcNative.check(ctx, (int) MfkcNative.SUBMIT_READS.invokeExact(ctx, hBases, hOffs, n));
                }
                MfkcNative.check(ctx, (int) MfkcNative.FLUSH.invokeExact(ctx));

                // IOUtils.printKmers :45-71
                MemorySegment nGood = arena.allocate(JAVA_LONG);
                MfkcNative.check(ctx, (int) MfkcNative.EMIT_BEGIN.invokeExact(ctx, threshold, nGood));
                MemorySegment written = arena.allocate(JAVA_LONG);
                byte[] chunk = new byte[16777200];                      // KMERS_WORK_RANGE_SIZE, IOUtils.java:30
                try (OutputStream out = new BufferedOutputStream(new FileOutputStream(outFile), 1 << 24)) {
                    while (true) {
                        MfkcNative.check(ctx, (int) MfkcNative.EMIT_NEXT.invokeExact(ctx, hBases, (long) chunk.length, written));
                        int w = (int) written.get(JAVA_LONG, 0);
                        if (w == 0) break;
                        MemorySegment.copy(hBases, JAVA_BYTE, 0, chunk, 0, w);
                        out.write(chunk, 0, w);
                    }
                }
                MemorySegment hist = arena.allocate(8L * 32768);
                MfkcNative.check(ctx, (int) MfkcNative.HISTOGRAM.invokeExact(ctx, hist));
                QuickQuantitativeStatistics<Short> stats = new QuickQuantitativeStatistics<Short>();
                for (int c = 1; c < 32768; c++) {
                    long v = hist.get(JAVA_LONG, 8L * c);
                    if (v != 0) stats.set((short) c, v);
                }
                stats.printToFile(stFile, "# k-mer frequency\tnumber of such k-mers");
                return nGood.get(JAVA_LONG, 0);
            } finally {
                MfkcNative.DESTROY.invokeExact(ctx);
            }
        } catch (ExecutionFailedException | IOException e) {
            throw e;
        } catch (Throwable t) {
            throw new ExecutionFailedException("libmfkc call failed", t);
        }
    }

    /**
     * The same ingest with the library's own readers (mfkc_reader_*: the parser rules of ReadersUtils / FastaReader /
     * FastqReader restated in C++, mapped input, parallel parse workers, a multi-threaded gzip decoder) instead of the ITMO
     * readers: 4-10 M reads/s instead of the single GZIPInputStream + parser thread of ReadsDispatcher.  BINQ / bz2 inputs
     * and IUPAC codes are not served by it (mfkc_reader_open reports MFKC_E_FORMAT): fall back to the loop above for those.
     * Returns the number of reads submitted.
     */
    static long submitFileNative(MemorySegment ctx, File file, MemorySegment hBases, long capBases, MemorySegment hOffs,
                                 int capReads, Arena arena) throws Throwable {
        MemorySegment pReader = arena.allocate(ADDRESS), err = arena.allocate(512), n = arena.allocate(JAVA_INT);
        int rc = (int) MfkcNative.READER_OPEN.invokeExact(arena.allocateFrom(file.getPath()), pReader, err, 512L);
        if (rc != 0) throw new ExecutionFailedException(err.getString(0));
        MemorySegment reader = pReader.get(ADDRESS, 0);
        long reads = 0;
        try {
            while (true) {
                rc = (int) MfkcNative.READER_NEXT.invokeExact(reader, hBases, capBases, hOffs, capReads, n);
                if (rc == -1) {                                    // MFKC_E_BADARG: a read longer than hBases stays pending
                    MemorySegment pending = arena.allocate(JAVA_LONG);
                    MfkcNative.READER_PENDING_BASES.invokeExact(reader, pending);
                    long need = pending.get(JAVA_LONG, 0);
                    if (need > capBases) {                         // (a chromosome-sized FASTA record: take it with a larger buffer)
                        MemorySegment pBuf = arena.allocate(ADDRESS);
                        MfkcNative.check(ctx, (int) MfkcNative.PINNED_ALLOC.invokeExact(ctx, need, pBuf));
                        hBases = pBuf.get(ADDRESS, 0).reinterpret(need);
                        capBases = need;
                        continue;
                    }
                }
                if (rc != 0) throw new ExecutionFailedException("Error while reading " + file.getName());   // text: mfkc_reader_error
                int got = n.get(JAVA_INT, 0);
                if (got == 0) break;
                MfkcNative.check(ctx, (int) MfkcNative.SUBMIT_READS.invokeExact(ctx, hBases, hOffs, got));
                reads += got;
            }
        } finally {
            MfkcNative.READER_CLOSE.invokeExact(reader);
        }
        return reads;
    }

    private GpuKmerCounting() {}
}
