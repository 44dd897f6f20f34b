package tools;

import io.MfkcNative;
import ru.ifmo.genetics.utils.tool.ExecutionFailedException;
import structures.ConnectedComponent;

import java.io.File;
import java.io.FileNotFoundException;
import java.io.PrintWriter;
import java.lang.foreign.Arena;
import java.lang.foreign.MemorySegment;
import java.nio.channels.FileChannel;
import java.nio.file.StandardOpenOption;
import java.util.List;

import static java.lang.foreign.ValueLayout.*;

/**
 * The GPU body of features-calculator: what a maintainer puts behind FeaturesCalculatorMain.runImpl
 * (src/tools/FeaturesCalculatorMain.java:77-166) -- same parameters, same output files, same log lines; only the map
 * (BigLong2LongHashMap, :97-103), the presence accumulation (IOUtils.calculatePresenceForKmers, src/io/IOUtils.java:577-597)
 * and the per-component sums (buildAndPrintVector, :169-236) run in libmfkc.  JDK >= 22 (java.lang.foreign, final API).
 * NOT compiled in this repository (no JDK in the build image); mfkc_cli's tool_features (metafast_b200/csrc/cli.cpp) is the
 * executable twin of this class and is what the tests run.
 */
public final class GpuFeaturesCalculatorMain {

    /** -ka files: one .vec and one .breadth per .kmers.bin file, components from -cm. */
    public static void run(int k, List<ConnectedComponent> components, File[] kmersFiles, long threshold, File outDir, int device)
            throws ExecutionFailedException {
        try (Arena arena = Arena.ofConfined()) {
            MemorySegment cfg = arena.allocate(MfkcNative.CFG);
            cfg.fill((byte) 0);
            cfg.set(JAVA_INT, 0, (int) MfkcNative.CFG.byteSize());
            cfg.set(JAVA_INT, 4, k);
            cfg.set(JAVA_INT, 12, device);
            MemorySegment pCtx = arena.allocate(ADDRESS);
            int rc = (int) MfkcNative.CREATE.invokeExact(cfg, pCtx);
            if (rc != 0) MfkcNative.check(MemorySegment.NULL, rc);
            MemorySegment ctx = pCtx.get(ADDRESS, 0);
            try {
                // hm.put(kmer, 0) for every component k-mer (:97-103): keys + CSR offsets
                long total = 0;
                for (ConnectedComponent c : components) total += c.kmers.size();
                MemorySegment keys = arena.allocate(JAVA_LONG, Math.max(total, 1));
                MemorySegment off = arena.allocate(JAVA_LONG, components.size() + 1L);
                long p = 0;
                for (int i = 0; i < components.size(); i++) {
                    off.setAtIndex(JAVA_LONG, i, p);
                    for (long kmer : components.get(i).kmers) keys.setAtIndex(JAVA_LONG, p++, kmer);
                }
                off.setAtIndex(JAVA_LONG, components.size(), p);
                MfkcNative.check(ctx, (int) MfkcNative.FC_LOAD.invokeExact(ctx, keys, off, components.size()));

                MemorySegment vec = arena.allocate(JAVA_LONG, components.size());
                MemorySegment found = arena.allocate(JAVA_LONG, components.size());
                MemorySegment cnt = arena.allocate(JAVA_LONG, components.size());
                for (File f : kmersFiles) {                                            // :136-163
                    MfkcNative.check(ctx, (int) MfkcNative.FC_RESET.invokeExact(ctx)); // hm.resetValues()
                    try (FileChannel ch = FileChannel.open(f.toPath(), StandardOpenOption.READ)) {
                        final long chunk = 16777200L;                                  // KMERS_WORK_RANGE_SIZE, src/io/IOUtils.java:30
                        for (long pos = 0; pos < ch.size(); pos += chunk) {
                            long n = Math.min(chunk, ch.size() - pos);
                            MemorySegment recs = ch.map(FileChannel.MapMode.READ_ONLY, pos, n, arena);
                            MfkcNative.check(ctx, (int) MfkcNative.FC_ADD_RECORDS.invokeExact(ctx, recs, n / 10));
                        }
                    }
                    MfkcNative.check(ctx, (int) MfkcNative.FC_FEATURES.invokeExact(ctx, threshold, vec, found, cnt));
                    String stem = f.getName().replaceAll("\\.kmers\\.bin$", "");
                    print(new File(outDir, stem + ".vec"), new File(outDir, stem + ".breadth"), vec, found, cnt, components.size());
                }
            } finally {
                MfkcNative.DESTROY.invokeExact(ctx);
            }
        } catch (ExecutionFailedException e) {
            throw e;
        } catch (Throwable t) {
            throw new ExecutionFailedException("libmfkc call failed", t);
        }
    }

    /** buildAndPrintVector's two files (:205-236): one long per line, one double (found / cnt) per line. */
    private static void print(File vecFile, File breadthFile, MemorySegment vec, MemorySegment found, MemorySegment cnt, int n)
            throws FileNotFoundException {
        try (PrintWriter v = new PrintWriter(vecFile); PrintWriter b = new PrintWriter(breadthFile)) {
            for (int i = 0; i < n; i++) {
                v.println(vec.getAtIndex(JAVA_LONG, i));
                b.println((double) found.getAtIndex(JAVA_LONG, i) / cnt.getAtIndex(JAVA_LONG, i));
            }
        }
    }

    private GpuFeaturesCalculatorMain() {}
}
