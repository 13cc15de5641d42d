"""Host-side mirror of the reference's plugin surface (src/index.js:16-31) over the C ABI.

The reference's host language (JavaScript on a WebGL renderer) has no toolchain in this image, so the
host side above the C ABI is written here in Python with the reference's class names, constructor
signatures, option names/defaults and methods; `js/` holds the same classes as ES modules over the
N-API shim for a box that has Node.  Every compute call below ends in one `rfx_*`
entry point of csrc/librfx.so — there is no CPU path.

What stands in for three.js / postprocessing objects:
  composer  : any object with `.ctx` (engine.Context), `.width/.height`, `.inputBuffer` (DevPlane RGBA16F, the
              scene colour = directLight) and `.outputBuffer` (DevPlane RGBA16F)
  scene     : any object with `.gbuffer`, `.depth` (packed G-buffer + depth planes, the output layout of
              GBufferPass, src/gbuffer/GBufferPass.js:33-44) and `.velocity` (VelocityDepthNormalPass layout)
              — rasterising meshes into those planes is out of scope
  camera    : synth.Camera or any object with `.uniforms()` (three.js matrices as float32 column-major)
"""
from __future__ import annotations

import math

import numpy as np

from . import abi, engine

# ---------------------------------------------------------------------------------------------------
# option tables (verbatim defaults)
# ---------------------------------------------------------------------------------------------------
defaultSSGIOptions = dict(  # src/ssgi/SSGIOptions.js:26-48
    mode="ssgi", distance=10, thickness=10, denoiseIterations=1, denoiseKernel=2, denoiseDiffuse=10, denoiseSpecular=10, radius=3, phi=0.5,
    lumaPhi=5, depthPhi=2, normalPhi=50, roughnessPhi=50, specularPhi=50, envBlur=0.5, importanceSampling=True, steps=20, refineSteps=5,
    resolutionScale=1, missedRays=False, outputTexture=None)
defaultTemporalReprojectPassOptions = dict(  # src/temporal-reproject/TemporalReprojectPass.js:17-32
    dilation=False, fullAccumulate=False, neighborhoodClamp=False, neighborhoodClampRadius=1, neighborhoodClampIntensity=1, maxBlend=1,
    logTransform=False, depthDistance=2, worldDistance=4, reprojectSpecular=False, renderTarget=None, copyTextures=True, confidencePower=0.75,
    inputType="diffuse")
defaultPoissonBlurOptions = dict(  # src/denoise/pass/PoissonDenoisePass.js:16-24
    iterations=1, radius=3, phi=0.5, lumaPhi=5, depthPhi=2, normalPhi=3.25, inputType="diffuseSpecular")
defaultAOOptions = dict(  # src/ao/AOEffect.js:8-21
    resolutionScale=1, spp=8, distance=2, distancePower=1, power=2, bias=40, thickness=0.075, color=(0.0, 0.0, 0.0), useNormalPass=False,
    velocityDepthNormalPass=None, normalTexture=None, **defaultPoissonBlurOptions)
defaultMotionBlurOptions = dict(intensity=1, jitter=1, samples=16)  # src/motion-blur/MotionBlurEffect.js:14

_HIGHEST_SIGNED_INT = 0x7FFFFFFF


class BlueNoiseIndex:
    """The `blueNoiseIndex` uniform of setupBlueNoise (src/utils/BlueNoiseUtils.js:17-33): advances on every read.
    `start` pins the reference's Math.random() seed."""

    def __init__(self, start: int = 1234567):
        self.start, self.i = int(start), 0

    @property
    def value(self) -> int:
        self.i = (self.start + self.i + 1) % _HIGHEST_SIGNED_INT
        return self.i


def _did_camera_move(cam_u: dict, last: dict | None) -> bool:
    """src/utils/SceneUtils.js:17-27 (position / orientation compared with small epsilons)"""
    if last is None:
        return True
    a, b = np.asarray(cam_u["camera_matrix_world"], np.float64), np.asarray(last["camera_matrix_world"], np.float64)
    return bool(np.abs(a[12:15] - b[12:15]).max() > 1e-6 or np.abs(a[:12] - b[:12]).max() > 1e-6)


class _Reactive:
    """Every option key becomes a get/set property of the effect (makeOptionsReactive)."""

    _options: dict

    def __getattr__(self, k):
        o = self.__dict__.get("_options")
        if o is not None and k in o:
            return o[k]
        raise AttributeError(k)

    def __setattr__(self, k, v):
        o = self.__dict__.get("_options")
        if o is not None and k in o and not k.startswith("_"):
            if o[k] != v:
                o[k] = v
                self._option_changed(k)
            return
        object.__setattr__(self, k, v)

    def _option_changed(self, k):  # pragma: no cover - overridden
        pass


# ---------------------------------------------------------------------------------------------------
class VelocityDepthNormalPass:
    """src/temporal-reproject/pass/VelocityDepthNormalPass.js:66-194.  The reference rasterises the scene into
    (uv motion, packed oct normal, depth); here the plane is supplied by the host (`scene.velocity`)."""

    needsSwap = False

    def __init__(self, scene, camera):
        self._scene, self._camera = scene, camera
        self.lastVelocityTexture = None

    @property
    def texture(self):
        return self._scene.velocity

    @property
    def renderTarget(self):
        return self

    @property
    def depthTexture(self):
        return self._scene.depth

    def setSize(self, width, height):
        pass

    def render(self, renderer=None):
        self.lastVelocityTexture = self._scene.velocity

    def dispose(self):
        pass


class VelocityPass(VelocityDepthNormalPass):
    """src/temporal-reproject/pass/VelocityPass.js:3-7: the same pass object (the host supplies the plane either way)"""


# ---------------------------------------------------------------------------------------------------
class SSGIEffect(_Reactive):
    """new SSGIEffect(composer, scene, camera, options)   (src/ssgi/SSGIEffect.js:27-141; signature per D6)

    update() runs K1 -> K2 -> K3 x 2*denoiseIterations -> K4 natively (rfx_ssgi_chain) and then K5 into
    composer.outputBuffer.  `outputTexture` is the composed GI plane (RGBA32F)."""

    DefaultOptions = defaultSSGIOptions

    def __init__(self, composer, scene, camera, options=None):
        opts = {**defaultSSGIOptions, **(options or {})}
        if opts["mode"] == "ssr":  # src/ssgi/SSGIEffect.js:70-77
            opts.update(reprojectSpecular=True, neighborhoodClamp=True, inputType="specular")
        else:
            opts.update(reprojectSpecular=[False, True], neighborhoodClamp=[False, True])
        opts.setdefault("denoiseMode", "full")  # Denoiser.js:6-11 (spread into the Denoiser through ...options, SSGIEffect.js:104-108)
        preset = opts.get("preset")
        if isinstance(preset, str):  # src/ssgi/SSGIEffect.js:79-99 (the third case is a second, unreachable "medium" in the reference)
            if preset == "low":
                opts.update(steps=10, refineSteps=2, denoiseMode="full_temporal")
            elif preset == "medium":
                opts.update(steps=20, refineSteps=4, denoiseMode="full")
        if opts["denoiseMode"] not in abi.DENOISE_MODES:
            raise abi.RfxError(f"denoiseMode {opts['denoiseMode']!r}: 'denoised' binds an array of textures to a sampler in the reference and cannot run there; "
                               f"supported: {sorted(abi.DENOISE_MODES)}")
        self.composer, self._scene, self._camera = composer, scene, camera
        self.ctx: engine.Context = composer.ctx
        self.velocityDepthNormalPass = opts.get("velocityDepthNormalPass") or VelocityDepthNormalPass(scene, camera)
        self.isUsingRenderPass = True  # directLight comes from the composer input buffer (useDirectLight)
        self._blue_start = int(opts.pop("blueNoiseStart", 1234567))
        self._last_cam = None
        self._chain = None
        self._options = opts
        self.setSize(composer.width, composer.height)

    # -- native chain ------------------------------------------------------------------------------
    def _chain_options(self) -> abi.ChainOptions:
        o, c = self._options, abi.ChainOptions()
        c.width, c.height = self._size
        c.denoise_iterations, c.steps, c.refine_steps = int(o["denoiseIterations"]), int(o["steps"]), int(o["refineSteps"])
        c.distance, c.thickness, c.env_blur = o["distance"], o["thickness"], o["envBlur"]
        c.radius, c.phi, c.luma_phi, c.depth_phi = o["radius"], o["phi"], o["lumaPhi"], o["depthPhi"]
        c.normal_phi, c.roughness_phi, c.specular_phi = o["normalPhi"], o["roughnessPhi"], o["specularPhi"]
        flags = 0
        if getattr(self.ctx, "has_env", False):
            flags |= abi.SSGI_USE_ENVMAP | (abi.SSGI_IMPORTANCE_SAMPLING if o["importanceSampling"] else 0)
        if o["missedRays"]:
            flags |= abi.SSGI_MISSED_RAYS
        if self.isUsingRenderPass:
            flags |= abi.SSGI_USE_DIRECT_LIGHT
        c.ssgi_flags = flags
        c.mode = abi.MODE_SSR if o["mode"] == "ssr" else abi.MODE_SSGI
        c.blue_noise_start = self._blue_start
        c.denoise_mode = abi.DENOISE_MODES[o["denoiseMode"]]
        c.resolution_scale = float(o["resolutionScale"])
        return c

    def setSize(self, width, height, force=False):
        if not force and getattr(self, "_size", None) == (width, height):
            return
        self._size = (int(width), int(height))
        if self._chain is not None:
            self._chain.close()
        self._chain = engine.SsgiChain(self.ctx, self._chain_options())

    def _option_changed(self, k):
        if k == "resolutionScale":
            self.setSize(*self._size, force=True)
        elif self._chain is not None:
            self._chain.set_options(self._chain_options())  # setters end with reset() in the reference

    def setEnvironment(self, map_f16, marginal=None, conditional=None, total_sum=1.0):
        """keepEnvMapUpdated (src/ssgi/SSGIEffect.js:309-366): equirect RGBA16F map + CDF tables"""
        self.ctx.set_env(map_f16, marginal, conditional, total_sum)
        self.ctx.has_env = True
        self._chain.set_options(self._chain_options())

    def reset(self):
        self._chain.reset()

    def initialize(self, renderer=None, *args):
        pass

    @property
    def depthTexture(self):
        return self._scene.depth

    @property
    def outputTexture(self):
        return self._chain.output(0)

    def update(self, renderer=None, inputBuffer=None, deltaTime=None):
        cam_u = self._camera.uniforms()
        moved = _did_camera_move(cam_u, self._last_cam)
        self._last_cam = cam_u
        self.velocityDepthNormalPass.render(renderer)
        scene_buf = inputBuffer if inputBuffer is not None else self.composer.inputBuffer
        self._chain.render(abi.make_camera(cam_u), self._scene.depth, self._scene.gbuffer, self.velocityDepthNormalPass.texture,
                           scene_buf if self.isUsingRenderPass else None, cam_u["position"], moved)
        if getattr(self.composer, "outputBuffer", None) is not None:  # both modes: SSR composes with inputType "specular" (Denoiser.js:56-64)
            self.ctx.ssgi_compose(self._scene.depth, self.outputTexture, scene_buf, self.composer.outputBuffer, params=self._compose_params(cam_u))  # K5 (mainImage of the effect)

    def _compose_params(self, cam_u):
        """scene.fog -> the fog uniforms of ssgi_compose.frag (src/ssgi/SSGIEffect.js:80-90, 404-412); fog = dict(color, near, far) or
        dict(color, density, isFogExp2=True) like three.js Fog / FogExp2"""
        fog = getattr(self._scene, "fog", None)
        if not fog:
            return None
        p = abi.SsgiComposeParams()
        p.use_fog, p.fog_exp2 = 1, int(bool(fog.get("isFogExp2", False)))
        p.fog_color[:] = [float(x) for x in fog.get("color", (1.0, 1.0, 1.0))]
        p.fog_near, p.fog_far, p.fog_density = float(fog.get("near", 1.0)), float(fog.get("far", 1000.0)), float(fog.get("density", 0.00025))
        p.camera_near, p.camera_far, p.perspective = float(cam_u["near"]), float(cam_u["far"]), 1
        return p

    def dispose(self):
        if self._chain is not None:
            self._chain.close()
            self._chain = None


class SSREffect(SSGIEffect):
    """src/ssgi/SSREffect.js:3-9"""

    def __init__(self, composer, scene, camera, options=None):
        super().__init__(composer, scene, camera, {**defaultSSGIOptions, **(options or {}), "mode": "ssr"})


# ---------------------------------------------------------------------------------------------------
def generateR2(count: int) -> list:
    """src/temporal-reproject/utils/QuasirandomGenerator.js:11-24 (JS doubles)"""
    g = 1.32471795724474602596090885447809  # plastic number
    a1, a2, base = 1.0 / g, 1.0 / (g * g), 1.1127756842787055
    return [[math.fmod(base + a1 * n, 1.0), math.fmod(base + a2 * n, 1.0)] for n in range(count)]


r2Sequence = [[a - 0.5, b - 0.5] for a, b in generateR2(256)]  # src/taa/TAAUtils.js:3


def jitter(width, height, camera, frame: int, jitterScale: float = 1.0):
    """src/taa/TAAUtils.js:5-11: sub-pixel view offset from the R2 sequence (a no-op for cameras without setViewOffset)"""
    x, y = r2Sequence[frame % len(r2Sequence)]
    if hasattr(camera, "setViewOffset"):
        camera.setViewOffset(width, height, x * jitterScale, y * jitterScale, width, height)


class TemporalReprojectPass:
    """new TemporalReprojectPass(scene, camera, velocityDepthNormalPass, texture, textureCount, options)
    (src/temporal-reproject/TemporalReprojectPass.js:38-225).  1-plane RGBA16F configuration (TRAA)."""

    needsSwap = False

    def __init__(self, scene, camera, velocityDepthNormalPass, texture, textureCount, options=None):
        if textureCount != 1:
            raise abi.RfxError("standalone TemporalReprojectPass supports textureCount == 1 (the 2-plane SSGI form runs inside SSGIEffect)")
        self._scene, self._camera, self.velocityDepthNormalPass = scene, camera, velocityDepthNormalPass
        self.options = {**defaultTemporalReprojectPassOptions, **(options or {})}
        self.inputTexture, self.textureCount = texture, textureCount
        self.ctx: engine.Context = texture.ctx
        self.keepData, self._prev, self.frame = 1.0, None, 0
        self.renderTarget = self.framebufferTexture = None
        self.setSize(texture.width, texture.height)

    def setSize(self, width, height):
        for p in (self.renderTarget, self.framebufferTexture):
            if p is not None:
                p.free()
        self.renderTarget = self.ctx.alloc(abi.FMT_RGBA16F, width, height)
        self.framebufferTexture = self.ctx.alloc(abi.FMT_RGBA16F, width, height)  # copyFramebufferToTexture history (:197-200)

    @property
    def texture(self):
        """renderTarget.texture[0]: the plane rendered by the most recent render() (TemporalReprojectPass.js:150-156)"""
        return self.renderTarget

    def reset(self):
        self.keepData = 0.0

    def jitter(self, jitterScale: float = 1.0):  # TemporalReprojectPass.js:216-220
        self.unjitter()
        jitter(self.renderTarget.width, self.renderTarget.height, self._camera, self.frame, jitterScale)

    def unjitter(self):  # :222-224
        if hasattr(self._camera, "clearViewOffset"):
            self._camera.clearViewOffset()

    def render(self, renderer=None):
        self.frame = (self.frame + 1) % 4096
        # the pass uploads the UN-jittered projection (camera.view disabled around updateProjectionMatrix, :168-186)
        cam_u = self._camera.unjittered_uniforms() if hasattr(self._camera, "unjittered_uniforms") else self._camera.uniforms()
        prev = self._prev or cam_u
        o, p = self.options, abi.TemporalParams()
        p.cam = abi.make_camera(cam_u)
        abi.set_f16(p.prev_view_matrix, prev["view_matrix"])
        abi.set_f16(p.prev_camera_matrix_world, prev["camera_matrix_world"])
        abi.set_f16(p.prev_projection, prev["projection"])
        abi.set_f16(p.prev_projection_inverse, prev["projection_inverse"])
        p.camera_pos[:] = [float(x) for x in cam_u["position"]]
        p.prev_camera_pos[:] = [float(x) for x in prev["position"]]
        p.max_blend, p.neighborhood_clamp_intensity = o["maxBlend"], o["neighborhoodClampIntensity"]
        p.keep_data, p.confidence_power = self.keepData, o["confidencePower"]
        p.full_accumulate = int(bool(o["fullAccumulate"]) and not _did_camera_move(cam_u, self._prev))
        p.texture_count, p.input_type, p.log_transform, p.history_linear = 1, abi.INPUT_DIFFUSE, int(bool(o["logTransform"])), 1
        p.reproject_specular[:] = [0, 0]
        # the reference copies the framebuffer into framebufferTexture after the draw (:197-200); here the two planes swap roles
        # BEFORE the launch instead, so that `renderTarget` / `texture` is always the plane this render() wrote
        self.renderTarget, self.framebufferTexture = self.framebufferTexture, self.renderTarget
        self.ctx.temporal_reproject(p, self.inputTexture, self.velocityDepthNormalPass.texture, self.framebufferTexture, None, self.renderTarget, None)
        self.keepData = 1.0
        self._prev = cam_u

    @property
    def accumulated(self):
        return self.renderTarget

    def dispose(self):
        for p in (self.renderTarget, self.framebufferTexture):
            if p is not None:
                p.free()


class TRAAEffect:
    """new TRAAEffect(scene, camera, velocityDepthNormalPass, options)  (src/traa/TRAAEffect.js:10-76)"""

    DefaultOptions = defaultTemporalReprojectPassOptions

    def __init__(self, scene, camera, velocityDepthNormalPass, options=None):
        self._scene, self._camera, self.velocityDepthNormalPass = scene, camera, velocityDepthNormalPass
        forced = dict(maxBlend=0.9, neighborhoodClamp=True, neighborhoodClampIntensity=1, neighborhoodClampRadius=1, logTransform=True, confidencePower=4)
        self.options = {**defaultTemporalReprojectPassOptions, **(options or {}), **forced}  # :21-33
        self.temporalReprojectPass = None

    def setSize(self, width, height):
        if self.temporalReprojectPass:
            self.temporalReprojectPass.setSize(width, height)

    def reset(self):
        self.temporalReprojectPass.reset()

    def update(self, renderer=None, inputBuffer=None, deltaTime=None):
        if self.temporalReprojectPass is None:
            self.temporalReprojectPass = TemporalReprojectPass(self._scene, self._camera, self.velocityDepthNormalPass, inputBuffer, 1, self.options)
        self.temporalReprojectPass.inputTexture = inputBuffer
        trp = self.temporalReprojectPass
        trp.unjitter()                                                       # TRAAEffect.js:67-72
        self.unjitteredProjectionMatrix = np.array(getattr(self._camera, "proj", np.eye(4)), copy=True)
        trp.jitter()                                                         # the NEXT scene render is rasterised with this sub-pixel offset
        trp.render(renderer)

    def compose(self, outputBuffer):
        """traa_compose.frag: rgb passthrough, alpha 1"""
        self.temporalReprojectPass.ctx.traa_compose(self.temporalReprojectPass.accumulated, outputBuffer)

    def dispose(self):
        if self.temporalReprojectPass:
            self.temporalReprojectPass.dispose()


# ---------------------------------------------------------------------------------------------------
class PoissonDenoisePass:
    """new PoissonDenoisePass(camera, textures, options)  (src/denoise/pass/PoissonDenoisePass.js:26-150)"""

    DefaultOptions = defaultPoissonBlurOptions

    def __init__(self, camera, textures, options=None):
        o = {**defaultPoissonBlurOptions, **(options or {})}
        self.options, self.textures, self.iterations = o, list(textures), o["iterations"]
        self.ctx: engine.Context = self.textures[0].ctx
        self.textureCount = 2 if o["inputType"] == "diffuseSpecular" else 1
        self.isTextureSpecular = {"diffuseSpecular": [0, 1], "diffuse": [0, 0], "specular": [1, 1]}[o["inputType"]]
        self.radius, self.phi, self.lumaPhi, self.depthPhi, self.normalPhi = o["radius"], o["phi"], o["lumaPhi"], o["depthPhi"], o["normalPhi"]
        self.roughnessPhi, self.specularPhi = o.get("roughnessPhi", 0.0), o.get("specularPhi", 0.0)
        self._gbuffer = self._depth = None
        self._is_gbuffer = False
        self.blueNoiseIndex = BlueNoiseIndex(o.get("blueNoiseStart", 1234567))
        self.renderTargetA = self.renderTargetB = None
        self.setSize(self.textures[0].width, self.textures[0].height)

    def setSize(self, width, height):
        self.dispose()
        self.renderTargetA = [self.ctx.alloc(abi.FMT_RGBA16F, width, height) for _ in range(self.textureCount)]
        self.renderTargetB = [self.ctx.alloc(abi.FMT_RGBA16F, width, height) for _ in range(self.textureCount)]

    @property
    def texture(self):
        return self.renderTargetB

    def setGBufferPass(self, gbuffer_plane, depth_plane, is_gbuffer=True):
        """GBufferPass => GBUFFER_TEXTURE; VelocityDepthNormalPass => velocity-layout normals (:109-118)"""
        self._gbuffer, self._depth, self._is_gbuffer = gbuffer_plane, depth_plane, bool(is_gbuffer)

    def render(self, renderer=None):
        for i in range(2 * self.iterations):
            horizontal = i % 2 == 0
            src = self.textures if i == 0 else (self.renderTargetB if horizontal else self.renderTargetA)
            dst = self.renderTargetA if horizontal else self.renderTargetB
            p = abi.PoissonParams()
            p.radius, p.phi, p.luma_phi, p.depth_phi, p.normal_phi = self.radius, self.phi, self.lumaPhi, self.depthPhi, self.normalPhi
            p.roughness_phi, p.specular_phi = self.roughnessPhi, self.specularPhi
            p.texture_count = self.textureCount
            p.is_texture_specular[:] = self.isTextureSpecular
            p.gbuffer_texture = int(self._is_gbuffer)
            p.input_linear = int(i > 0 or src[0].format == abi.FMT_RGBA16F)
            p.blue_noise_index = self.blueNoiseIndex.value
            two = self.textureCount == 2
            self.ctx.poisson_denoise(p, self._depth, self._gbuffer, src[0], src[1] if two else None, dst[0], dst[1] if two else None)

    def dispose(self):
        for rt in (self.renderTargetA, self.renderTargetB):
            for p in rt or []:
                p.free()
        self.renderTargetA = self.renderTargetB = None


# ---------------------------------------------------------------------------------------------------
class HBAOEffect(_Reactive):
    """new HBAOEffect(composer, camera, scene, options)  (src/hbao/HBAOEffect.js:5-20, src/ao/AOEffect.js:23-178).
    The reference class does not compile at its pinned commit; this is hbao.frag + the 1-plane
    velocity-layout Poisson denoise + ao_compose.frag, the wiring the shaders are written for."""

    DefaultOptions = defaultAOOptions

    def __init__(self, composer, camera, scene, options=None):
        self.composer, self._camera, self._scene = composer, camera, scene
        self.ctx: engine.Context = composer.ctx
        self.blueNoiseIndex = BlueNoiseIndex((options or {}).get("blueNoiseStart", 1234567))
        o = {**defaultAOOptions, **(options or {})}
        o.pop("blueNoiseStart", None)
        self.aoTarget = self.ctx.alloc(abi.FMT_RGBA16F, composer.width, composer.height)
        self.PoissonDenoisePass = PoissonDenoisePass(camera, [self.aoTarget], dict(iterations=o["iterations"], radius=o["radius"], phi=o["phi"],
                                                                                    lumaPhi=o["lumaPhi"], depthPhi=o["depthPhi"],
                                                                                    normalPhi=o["normalPhi"], inputType="diffuse"))
        self._options = o

    def _option_changed(self, k):
        if k in ("iterations", "radius", "phi"):
            setattr(self.PoissonDenoisePass, k, self._options[k])
        elif k in ("lumaPhi", "depthPhi", "normalPhi"):
            setattr(self.PoissonDenoisePass, k, max(self._options[k], 0.0001))  # AOEffect.js:107-111

    @property
    def texture(self):
        return self.PoissonDenoisePass.texture[0] if self._options["iterations"] > 0 else self.aoTarget

    def update(self, renderer=None, inputBuffer=None, deltaTime=None):
        o, cam_u = self._options, self._camera.uniforms()
        p = abi.HbaoParams()
        P = np.asarray(cam_u["projection"], np.float64).reshape(4, 4).T
        V = np.asarray(cam_u["view_matrix"], np.float64).reshape(4, 4).T
        abi.set_f16(p.projection_view, np.ascontiguousarray((P @ V).T.reshape(16)).astype(np.float32))  # AOPass.js:93-96
        abi.set_f16(p.projection_inverse, cam_u["projection_inverse"])
        abi.set_f16(p.camera_matrix_world, cam_u["camera_matrix_world"])
        p.ao_distance, p.distance_power, p.bias, p.thickness = o["distance"], o["distancePower"], o["bias"], o["thickness"]
        p.spp, p.blue_noise_index = int(o["spp"]), self.blueNoiseIndex.value
        self.ctx.hbao(p, self._scene.depth, self.aoTarget)
        self.PoissonDenoisePass.setGBufferPass(self._scene.velocity, self._scene.depth, is_gbuffer=False)
        self.PoissonDenoisePass.render(renderer)
        if inputBuffer is not None and getattr(self.composer, "outputBuffer", None) is not None:
            c = abi.AoComposeParams()
            c.power = o["power"]
            c.color[:] = [float(x) for x in o["color"]]
            self.ctx.ao_compose(c, self._scene.depth, self.texture, inputBuffer, self.composer.outputBuffer)

    def dispose(self):
        self.PoissonDenoisePass.dispose()
        self.aoTarget.free()


# ---------------------------------------------------------------------------------------------------
class MotionBlurEffect(_Reactive):
    """new MotionBlurEffect(velocityPass, options)  (src/motion-blur/MotionBlurEffect.js:16-102)"""

    def __init__(self, velocityPass, options=None):
        self.velocityPass = velocityPass
        self._frame = 0
        self._options = {**defaultMotionBlurOptions, **(options or {})}

    def update(self, renderer=None, inputBuffer=None, deltaTime=1 / 60, outputBuffer=None, window=None):
        """window = (innerWidth, innerHeight): the reference feeds the CSS window size as `resolution` (A10)."""
        ctx: engine.Context = inputBuffer.ctx
        p = abi.MotionBlurParams()
        p.intensity, p.jitter = self._options["intensity"], self._options["jitter"]
        p.delta_time = max(1 / 1000, deltaTime)
        p.resolution[:] = list(window or (inputBuffer.width, inputBuffer.height))
        p.frame = self._frame % 4096  # renderer.info.render.frame % 4096
        p.samples = int(self._options["samples"])
        self._frame += 1
        ctx.motion_blur(p, self.velocityPass.texture, inputBuffer, outputBuffer)


def getMaxMipLevel(width: int, height: int) -> int:
    """src/ssgi/utils/Utils.js:30-34"""
    return math.floor(math.log2(max(width, height))) + 1


# ---------------------------------------------------------------------------------------------------
# Cosmetic effects of the plugin surface (src/index.js:25-31) and the pass that merges them.
# ---------------------------------------------------------------------------------------------------
class SharpnessEffect:
    """new SharpnessEffect(options)  (src/sharpness/SharpnessEffect.js:36-59)"""

    fx_id = abi.FX_SHARPNESS

    def __init__(self, options=None):
        self.sharpness = {"sharpness": 1, **(options or {})}["sharpness"]

    def setSharpness(self, sharpness):
        self.sharpness = sharpness

    def update(self, renderer=None, inputBuffer=None, deltaTime=None):
        pass  # `inputTexture = inputBuffer.texture`: the merged kernel reads the pass input directly

    def _fill(self, p: abi.EffectsParams):
        p.sharpness = float(self.sharpness)


class LensDistortionEffect:
    """new LensDistortionEffect({ alphax, alphay, aberration })  (src/lens-distortion/LensDistortionEffect.js:48-77)"""

    fx_id = abi.FX_LENS_DISTORTION

    def __init__(self, options=None):
        o = {"alphax": -0.05, "alphay": -0.05, "aberration": 1, **(options or {})}
        self.alphax, self.alphay, self.aberration = o["alphax"], o["alphay"], o["aberration"]

    def setAlphaX(self, value):
        self.alphax = value

    def setAlphaY(self, value):
        self.alphay = value

    def update(self, renderer=None, inputBuffer=None, deltaTime=None):
        pass

    def _fill(self, p: abi.EffectsParams):
        p.alphax, p.alphay, p.aberration = float(self.alphax), float(self.alphay), float(self.aberration)


class GradualBackgroundEffect:
    """new GradualBackgroundEffect(camera, depthTexture, backgroundColor, maxDistance = 5)  (src/gradual-background/GradualBackgroundEffect.js:48-70)"""

    fx_id = abi.FX_GRADUAL_BACKGROUND

    def __init__(self, camera, depthTexture, backgroundColor, maxDistance=5):
        self._camera, self.depthTexture, self.backgroundColor, self.maxDistance = camera, depthTexture, tuple(backgroundColor), maxDistance

    def setBackgroundColor(self, color):
        self.backgroundColor = tuple(color)

    def setMaxDistance(self, distance):
        self.maxDistance = distance

    def update(self, renderer=None, inputBuffer=None, deltaTime=None):
        pass

    def _fill(self, p: abi.EffectsParams):
        p.background_color[:] = [float(c) for c in self.backgroundColor]
        p.max_distance = float(self.maxDistance)


class SparkleEffect:
    """new SparkleEffect(camera, velocityDepthNormalPass)  (src/sparkle/SparkleEffect.js:102-136).  The reference never defines
    PERSPECTIVE_CAMERA for this effect, so its getViewZ takes the orthographic branch; `definePerspectiveCamera = True` gives what a host
    that defines it gets."""

    fx_id = abi.FX_SPARKLE

    def __init__(self, camera, velocityDepthNormalPass, definePerspectiveCamera=False):
        self._camera, self.velocityDepthNormalPass = camera, velocityDepthNormalPass
        self.spread, self.intensity, self.definePerspectiveCamera = 1, 1, definePerspectiveCamera

    def setSpread(self, spread):
        self.spread = spread

    def setIntensity(self, intensity):
        self.intensity = intensity

    def update(self, renderer=None, inputBuffer=None, deltaTime=None):
        pass

    def _fill(self, p: abi.EffectsParams):
        p.spread, p.intensity, p.sparkle_perspective = float(self.spread), float(self.intensity), int(bool(self.definePerspectiveCamera))


class EffectPass:
    """postprocessing's `new EffectPass(camera, ...effects)` for the four effects above: the effects of one pass are merged into one
    fullscreen program — here ONE launch of rfx_effects_launch (csrc/k_fx.cu) — in which every effect samples the pass's input buffer
    and the colour flows from one effect to the next in the given order."""

    def __init__(self, camera, *effects):
        if not 1 <= len(effects) <= 4:
            raise ValueError("EffectPass: 1..4 effects")
        self._camera, self.effects = camera, list(effects)

    def render(self, renderer, inputBuffer, outputBuffer, deltaTime=None, stencilTest=None):
        ctx: engine.Context = inputBuffer.ctx
        p = abi.make_effects_params(self._camera.uniforms(), [e.fx_id for e in self.effects])
        depth = velocity = None
        for e in self.effects:
            e.update(renderer, inputBuffer, deltaTime)
            e._fill(p)
            if isinstance(e, GradualBackgroundEffect):
                depth = e.depthTexture
            if isinstance(e, SparkleEffect):
                velocity = e.velocityDepthNormalPass.texture
        ctx.effects(p, inputBuffer, depth, velocity, outputBuffer)


class TAAPass:
    """new TAAPass(camera)  (src/taa/TAAPass.js:18-95): the still-camera accumulator that renders to the screen.  `canvas` (RGBA8) stands in
    for the default framebuffer; the FramebufferTexture copy of it is the history.  `srgbOutput` = the renderer's output colour space."""

    renderToScreen = True

    def __init__(self, camera, srgbOutput=True):
        self._camera, self.srgbOutput = camera, srgbOutput
        self.cameraNotMovedFrames, self.frame, self.needsUpdate = 0, 0, False
        self._last = None
        self.canvas = self.framebufferTexture = None

    def setSize(self, width, height, ctx: "engine.Context | None" = None):
        """setSize(width, height) as in the reference (:57-66); the planes are allocated on `ctx`, or on the input buffer's context at the next render()"""
        self.dispose()
        self._size, self.needsUpdate = (int(width), int(height)), True
        if ctx is not None:
            self.canvas = ctx.alloc(abi.FMT_RGBA8, width, height)
            self.framebufferTexture = ctx.alloc(abi.FMT_RGBA8, width, height)

    def render(self, renderer, inputBuffer):
        ctx: engine.Context = inputBuffer.ctx
        if self.canvas is None:
            w, h = getattr(self, "_size", (inputBuffer.width, inputBuffer.height))
            self.canvas, self.framebufferTexture = ctx.alloc(abi.FMT_RGBA8, w, h), ctx.alloc(abi.FMT_RGBA8, w, h)
        self.frame = (self.frame + 1) % 4096
        cam_u = self._camera.uniforms()
        moved = self.needsUpdate or _did_camera_move(cam_u, self._last)
        self.needsUpdate = False
        n = self.cameraNotMovedFrames
        if n > 0:  # :81-84
            jitter(self.canvas.p.width, self.canvas.p.height, self._camera, self.frame, 1)
        self.cameraNotMovedFrames = 0 if moved else (n + 1) % 4096
        self._last = cam_u
        p = abi.TaaParams()
        p.camera_not_moved_frames, p.srgb_output = float(self.cameraNotMovedFrames), int(bool(self.srgbOutput))
        ctx.taa(p, inputBuffer, self.framebufferTexture, self.canvas)
        self.framebufferTexture, self.canvas = self.canvas, self.framebufferTexture  # copyFramebufferToTexture (:93): the canvas becomes the history
        return self.framebufferTexture

    def dispose(self):
        for pl in (self.canvas, self.framebufferTexture):
            if pl is not None:
                pl.free()
        self.canvas = self.framebufferTexture = None
