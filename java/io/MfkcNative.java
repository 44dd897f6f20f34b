package io;

import java.lang.foreign.*;
import java.lang.invoke.MethodHandle;

import static java.lang.foreign.ValueLayout.*;

/**
 * Panama FFM (JDK >= 22) binding of libmfkc.so -- include/mfkc.h, one downcall handle per entry point.
 * NOT compiled in this repository (no JDK in the build image); it is the binding a MetaFast maintainer
 * adds next to src/io/IOUtils.java.  Every method is a 1:1 image of the C prototype.
 */
public final class MfkcNative {
    private static final Linker LINKER = Linker.nativeLinker();
    private static final SymbolLookup LIB = SymbolLookup.libraryLookup(
            System.getProperty("mfkc.library", "libmfkc.so"), Arena.global());

    private static MethodHandle h(String name, FunctionDescriptor fd) {
        return LINKER.downcallHandle(LIB.find(name).orElseThrow(), fd);
    }

    /** struct mfkc_cfg (88 bytes, see include/mfkc.h) */
    public static final StructLayout CFG = MemoryLayout.structLayout(
            JAVA_INT.withName("struct_size"), JAVA_INT.withName("k"), JAVA_INT.withName("min_seq_len"),
            JAVA_INT.withName("device"), JAVA_INT.withName("variant"), JAVA_INT.withName("n_shards"),
            JAVA_INT.withName("shard_id"), JAVA_INT.withName("reserved0"),
            JAVA_LONG.withName("table_slots"), JAVA_LONG.withName("expected_distinct"),
            JAVA_LONG.withName("max_table_bytes"), JAVA_LONG.withName("staging_bytes"),
            JAVA_INT.withName("region_shift"), JAVA_INT.withName("reserved2"),
            JAVA_LONG.withName("expected_kmers"), JAVA_LONG.withName("reserved1"));

    static final MethodHandle CREATE = h("mfkc_create", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
    static final MethodHandle DESTROY = h("mfkc_destroy", FunctionDescriptor.ofVoid(ADDRESS));
    static final MethodHandle LAST_ERROR = h("mfkc_last_error", FunctionDescriptor.of(ADDRESS, ADDRESS));
    static final MethodHandle RESET = h("mfkc_reset", FunctionDescriptor.of(JAVA_INT, ADDRESS));
    static final MethodHandle PINNED_ALLOC = h("mfkc_pinned_alloc", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_LONG, ADDRESS));
    static final MethodHandle SUBMIT_READS = h("mfkc_submit_reads", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, JAVA_INT));
    static final MethodHandle FLUSH = h("mfkc_flush", FunctionDescriptor.of(JAVA_INT, ADDRESS));
    static final MethodHandle STATS = h("mfkc_stats", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
    static final MethodHandle HISTOGRAM = h("mfkc_histogram", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
    static final MethodHandle EMIT_BEGIN = h("mfkc_emit_begin", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS));
    static final MethodHandle EMIT_NEXT = h("mfkc_emit_next", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG, ADDRESS));
    static final MethodHandle FC_LOAD = h("mfkc_fc_load_components", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, JAVA_INT));
    static final MethodHandle FC_SELECTED = h("mfkc_fc_set_selected", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG));
    static final MethodHandle FC_RESET = h("mfkc_fc_reset_values", FunctionDescriptor.of(JAVA_INT, ADDRESS));
    static final MethodHandle FC_ADD_RECORDS = h("mfkc_fc_add_records", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG));
    static final MethodHandle FC_ADD_READS = h("mfkc_fc_add_reads", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, JAVA_INT));
    static final MethodHandle FC_FEATURES = h("mfkc_fc_features", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_LONG, ADDRESS, ADDRESS, ADDRESS));
    static final MethodHandle READER_OPEN = h("mfkc_reader_open", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, JAVA_LONG));
    static final MethodHandle READER_NEXT = h("mfkc_reader_next", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG, ADDRESS, JAVA_INT, ADDRESS));
    static final MethodHandle READER_CLOSE = h("mfkc_reader_close", FunctionDescriptor.ofVoid(ADDRESS));

    /** rc != 0 -> ExecutionFailedException with mfkc_last_error(ctx), like the tools' IOException wrapping (Tool.java:286-287). */
    static void check(MemorySegment ctx, int rc) throws ru.ifmo.genetics.utils.tool.ExecutionFailedException {
        if (rc == 0) return;
        try {
            MemorySegment msg = (MemorySegment) LAST_ERROR.invokeExact(ctx);
            throw new ru.ifmo.genetics.utils.tool.ExecutionFailedException(
                    "libmfkc error " + rc + ": " + msg.reinterpret(4096).getString(0));
        } catch (ru.ifmo.genetics.utils.tool.ExecutionFailedException e) {
            throw e;
        } catch (Throwable t) {
            throw new ru.ifmo.genetics.utils.tool.ExecutionFailedException("libmfkc error " + rc, t);
        }
    }

    private MfkcNative() {}
}
